// libvqb200.so -- the C ABI declared in include/vqb200.h.  Host-side launch logic only; the kernels live in
// the k*_*.cuh headers.  Built for sm_100a only (see build.py); there is no CPU path.
#include "vq_common.cuh"
#include "k1_prepare.cuh"
#include "k1_assign_simt.cuh"
#include "k1_assign_tc.cuh"
#include "k2_gather.cuh"
#include "k2_fused_ema.cuh"
#include "k3_ema.cuh"
#include "k3_p2p.cuh"

#include <algorithm>
#include <cstdlib>
#include <new>

using namespace vq;

static int check_shape(int64_t N, int64_t D, int64_t T, int K) {
    VQ_REQUIRE(N >= 0 && T >= 0, "negative batch or length");
    VQ_REQUIRE(D > 0 && D <= 65536, "emb_width out of range");
    VQ_REQUIRE(K > 0, "k_bins must be positive");
    VQ_REQUIRE(N * T < (int64_t(1) << 31), "more than 2^31 frames in one call");
    return 0;
}

static int gather_smem_bytes(int D, int& Ds, bool vec) {
    Ds = (D % 2 == 0) ? D + 1 : D;
    const size_t es = vec ? size_t(D) * (G_TT + 4) : size_t(G_TT) * Ds;
    return int(((es + 3) & ~size_t(3)) * 4 + G_TT * 8);
}

template <int MODE, bool VEC>
static int launch_gather_v(const float* x, const int64_t* idx, const float* mask, const float* k, const float* grad_xq,
                           const float* grad_commit, int64_t N, int D, int64_t T, int K, float* out, double* scalars,
                           float* results, cudaStream_t stream) {
    int Ds;
    int smem = gather_smem_bytes(D, Ds, VEC);
    VQ_REQUIRE(smem <= 200 * 1024, "emb_width too large for the gather tile (max ~750)");
    VQ_CUDA_OK(ensure_dynamic_smem(gather_kernel<MODE, VEC>, 200 * 1024));
    int64_t tiles = N * ((T + G_TT - 1) / G_TT);
    int per_sm = std::max(1, std::min(8, (220 * 1024) / (smem + 1024)));
    int grid = int(std::min<int64_t>(tiles, int64_t(num_sms()) * per_sm));
    const int dec4 = (MODE == GM_DECODE && T % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) ? 1 : 0;
    gather_kernel<MODE, VEC><<<grid, G_THREADS, smem, stream>>>(x, idx, mask, k, grad_xq, grad_commit, N, D, T, K, Ds, out,
                                                                scalars, results, (unsigned int)grid, dec4);
    VQ_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int MODE>
static int launch_gather_async(const float* x, const int64_t* idx, const float* mask, const float* k, const float* grad_xq,
                               const float* grad_commit, int64_t N, int D, int64_t T, int K, float* out, double* scalars,
                               float* results, cudaStream_t stream) {
    VQ_CUDA_OK(ensure_dynamic_smem(gather_async_kernel<MODE>, int(GaCfg<MODE>::SMEM)));
    const int64_t units = N * ((T + G_TT - 1) / G_TT) * ((D + GA_DS - 1) / GA_DS);
    const int grid = int(std::min<int64_t>(units, int64_t(num_sms()) * GaCfg<MODE>::CTAS_PER_SM));
    gather_async_kernel<MODE><<<grid, GA_THREADS, GaCfg<MODE>::SMEM, stream>>>(x, idx, mask, k, grad_xq, grad_commit, int(N), D, int(T), K,
                                                                               out, scalars, results, (unsigned int)grid);
    VQ_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int MODE>
static int launch_gather(const float* x, const int64_t* idx, const float* mask, const float* k, const float* grad_xq,
                         const float* grad_commit, int64_t N, int D, int64_t T, int K, float* out, double* scalars,
                         float* results, cudaStream_t stream) {
    auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    // measured on B200 (K=512, D=128): 16-byte accesses help the backward (3 streams) and decode, not the forward
    // measured on B200 (K=512, D=128): the 16-byte path wins for the backward (three streams), the 4-byte path elsewhere
    bool vec = MODE == GM_BWD && (T % 4 == 0) && aligned(x) && aligned(out) && aligned(grad_xq) && aligned(mask);
    static const char* force = getenv("VQ_K2_VEC");          // A/B switch for experiments: 0 = 4-byte path, 1 = 16-byte path
    if (force && (T % 4 == 0) && aligned(x) && aligned(out) && aligned(grad_xq) && aligned(mask)) vec = force[0] == '1';
    // default: the asynchronous persistent kernel whenever 16-byte copies are legal
    static const char* sync_only = getenv("VQ_K2_SYNC");       // A/B switch: 1 = the synchronous kernels below
    const int64_t units = N * ((T + G_TT - 1) / G_TT) * ((D + GA_DS - 1) / GA_DS);
    // (decode is a pure 227 MB write: the synchronous kernel with its L1-resident codebook measured 0.060 ms against 0.066)
    if (MODE != GM_DECODE && !(sync_only && sync_only[0] == '1') && (T % 4 == 0) && aligned(x) && aligned(out) && aligned(grad_xq) &&
        (mask == nullptr || true) && units < (int64_t(1) << 31) && T < (int64_t(1) << 31))
        return launch_gather_async<MODE>(x, idx, mask, k, grad_xq, grad_commit, N, D, T, K, out, scalars, results, stream);
    if (vec) return launch_gather_v<MODE, true>(x, idx, mask, k, grad_xq, grad_commit, N, D, T, K, out, scalars, results, stream);
    return launch_gather_v<MODE, false>(x, idx, mask, k, grad_xq, grad_commit, N, D, T, K, out, scalars, results, stream);
}

// int64 indices [nn, T] -> ragged uint16 codes: utterance n keeps its first len[n] frames at out[off[n] ..] (what
// dump_batch_to_pickle's q[:ql] keeps, scripts/generate_vq_dataset.py:83-90) -- 2 bytes per valid frame cross PCIe instead of 8 per padded one.
__global__ void __launch_bounds__(256) pack_codes_u16_kernel(const int64_t* __restrict__ idx, int64_t T, const int64_t* __restrict__ off,
                                                            unsigned short* __restrict__ out) {
    const int64_t n = blockIdx.y;
    const int64_t len = off[n + 1] - off[n];
    for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < len; t += int64_t(gridDim.x) * blockDim.x)
        out[off[n] + t] = static_cast<unsigned short>(idx[n * T + t]);
}

// ---- optional per-kernel event timing of vq_assign (bench.py roofline)
namespace {
constexpr int PROF_RING = 64;
struct Profiler {
    bool on = false;
    int n = 0;
    cudaEvent_t ev[PROF_RING][4];
    bool created = false;
} g_prof;
inline void prof_mark(int slot, int which, cudaStream_t s) {
    if (g_prof.on && slot >= 0) cudaEventRecord(g_prof.ev[slot][which], s);
}
}  // namespace

extern "C" {

int vq_profile_enable(int on) {
    if (on && !g_prof.created) {
        for (int i = 0; i < PROF_RING; ++i)
            for (int j = 0; j < 4; ++j) VQ_CUDA_OK(cudaEventCreate(&g_prof.ev[i][j]));
        g_prof.created = true;
    }
    g_prof.on = on != 0;
    g_prof.n = 0;
    return 0;
}

int vq_profile_read(float* ms4) {
    VQ_REQUIRE(ms4, "null pointer");
    ms4[0] = ms4[1] = ms4[2] = ms4[3] = 0.f;
    if (!g_prof.created || g_prof.n == 0) return 1;
    const int n = g_prof.n;
    VQ_CUDA_OK(cudaEventSynchronize(g_prof.ev[n - 1][3]));
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < 3; ++j) {
            float ms = 0.f;
            VQ_CUDA_OK(cudaEventElapsedTime(&ms, g_prof.ev[i][j], g_prof.ev[i][j + 1]));
            ms4[j] += ms / n;
        }
    ms4[3] = float(n);
    return 0;
}

int vq_version(void) { return VQB200_VERSION; }

const char* vq_last_error(void) { return err_buf(); }

int vq_device_supported(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10 ? 1 : 0;
}

size_t vq_workspace_bytes(int64_t n_utt, int64_t t_frames, int k_bins, int emb_width) {
    if (n_utt < 0 || t_frames < 0 || k_bins <= 0 || emb_width <= 0) return 0;
    AssignWorkspace w = carve_workspace(nullptr, n_utt * t_frames, k_bins, emb_width);
    return w.bytes + 256;
}

extern "C++" {
template <typename XT>
static int assign_impl(const XT* x, int64_t N, int64_t D, int64_t T, const float* k, int K,
                       int64_t* idx, float* min_d, double* scalars, void* workspace, size_t workspace_bytes,
                       int algo, void* stream_, float* dbg, long long* trace = nullptr, int trace_tiles = 0) {
    if (check_shape(N, D, T, K)) return 1;
    if (N * T == 0) return 0;
    VQ_REQUIRE(x && k && idx && workspace, "null pointer");
    VQ_REQUIRE(vq_device_supported(), "this library only runs on compute capability 10.x (B200); no fallback exists");
    VQ_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    AssignWorkspace w = carve_workspace(workspace, N * T, K, int(D));
    VQ_REQUIRE(workspace_bytes >= w.bytes, "workspace too small (see vq_workspace_bytes)");

    const bool prepared = (algo & VQ_ALGO_PREPARED) != 0;
    algo &= ~VQ_ALGO_PREPARED;
    bool use_tc = false;
    if (algo == VQ_ALGO_TC) {
        const char* why = tc_unsupported_reason(x, N, int(D), T, K, int(sizeof(XT)));
        if (why) return fail("vq_assign: VQ_ALGO_TC requested but %s", why);
        use_tc = true;
    } else if (algo == VQ_ALGO_AUTO) {
        use_tc = tc_unsupported_reason(x, N, int(D), T, K, int(sizeof(XT))) == nullptr;
    } else {
        VQ_REQUIRE(algo == VQ_ALGO_SIMT, "unknown algo");
    }

    const int pslot = (g_prof.on && g_prof.n < PROF_RING) ? g_prof.n++ : -1;
    prof_mark(pslot, 0, stream);
    if (!prepared) {
        VQ_CUDA_OK(cudaMemsetAsync(w.hdr, 0, sizeof(AssignHeader), stream));
        int blocks = (w.Kp * 32 + 255) / 256;
        codebook_prepare_kernel<<<blocks, 256, 0, stream>>>(k, K, int(D), w.Kp, w.Dp, w.ee, w.hn,
                                                            w.eb, w.hdr);
        VQ_CUDA_OK(cudaGetLastError());
    }
    prof_mark(pslot, 1, stream);
    if (use_tc) {
        if (launch_assign_tc(x, N, int(D), T, k, K, idx, min_d, scalars, w, stream, dbg, trace, trace_tiles,
                             trace ? (dbg != nullptr) : -1)) return 1;
        prof_mark(pslot, 2, stream);
        // exact re-scan of the frames whose FP16 shortlist could not be proven safe (the count lives on the device), one
        // warp per frame over the codes the tcgen05 pass could not rule out.
        // Programmatic dependent launch: the kernel is set up while the main kernel still runs (which signals
        // launch_dependents right after its prologue) and waits on griddepcontrol.wait before it reads the count, so the
        // usual empty-worklist case costs ~3 us less than a serialised launch.
        const int gx = int(std::min<int64_t>(4 * int64_t(num_sms()), (N * T + L_WARPS - 1) / L_WARPS));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(gx);
        cfg.blockDim = dim3(L_WARPS * 32);
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = (g_prof.on && pslot >= 0) ? 0 : 1;          // (event records between the two kernels serialise them anyway)
        const bool vec = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(k) & 15) == 0);
        const int n_code_tiles = w.Kp / 128;
        int own_shift = 0;                                     // how the main kernel assigned code tiles to its two scan groups
        {
            tc::Params plan;
            size_t plan_smem = 0;
            if (plan_assign_tc(int(D), K, plan, plan_smem) == nullptr) own_shift = plan.own_shift;
        }
        int dev = 0;
        VQ_CUDA_OK(cudaGetDevice(&dev));
        unsigned int* hint_dev = hard_hint(dev);               // (unified addressing: the mapped host pointer is valid on the device)
        auto launch_list = [&](auto kernel) {
            return cudaLaunchKernelEx(&cfg, kernel, x, N, int(D), T, k, (const float*)w.ee, K, n_code_tiles, own_shift, idx, min_d, scalars,
                                      (const int*)w.unsafe_rows, (const uint32_t*)w.unsafe_mask, (const uint2*)w.unsafe_tiles,
                                      tc::tile_scale_for(n_code_tiles), w.hdr, hint_dev);
        };
        if (vec && D <= 128) VQ_CUDA_OK(launch_list(assign_list_kernel<true, 1, XT>));
        else if (vec) VQ_CUDA_OK(launch_list(assign_list_kernel<true, 4, XT>));
        else if (D <= 128) VQ_CUDA_OK(launch_list(assign_list_kernel<false, 1, XT>));
        else VQ_CUDA_OK(launch_list(assign_list_kernel<false, 4, XT>));
    } else {
        int64_t tiles = N * ((T + S_BM - 1) / S_BM);
        int grid = int(std::min<int64_t>(tiles, int64_t(num_sms()) * 16));
        assign_simt_kernel<XT><<<grid, 256, 0, stream>>>(x, N, int(D), T, k, w.ee, K, idx, min_d, scalars);
        VQ_CUDA_OK(cudaGetLastError());
        prof_mark(pslot, 2, stream);
    }
    prof_mark(pslot, 3, stream);
    return 0;
}
}  // extern "C++"

int vq_assign(const float* x, int64_t N, int64_t D, int64_t T, const float* k, int K,
              int64_t* idx, float* min_d, double* scalars, void* workspace, size_t workspace_bytes,
              int algo, void* stream) {
    return assign_impl(x, N, D, T, k, K, idx, min_d, scalars, workspace, workspace_bytes, algo, stream, nullptr);
}

int vq_assign_bf16(const void* x_bf16, int64_t N, int64_t D, int64_t T, const float* k, int K,
                   int64_t* idx, float* min_d, double* scalars, void* workspace, size_t workspace_bytes,
                   int algo, void* stream) {
    return assign_impl(static_cast<const __nv_bfloat16*>(x_bf16), N, D, T, k, K, idx, min_d, scalars, workspace, workspace_bytes, algo, stream,
                       nullptr);
}

int vq_assign_debug(const float* x, int64_t N, int64_t D, int64_t T, const float* k, int K,
                    int64_t* idx, float* shortlist4, double* scalars, void* workspace, size_t workspace_bytes, void* stream,
                    int64_t* trace, int trace_tiles) {
    VQ_REQUIRE(shortlist4 || trace, "nothing to record");
    return assign_impl(x, N, D, T, k, K, idx, nullptr, trace && !shortlist4 ? nullptr : scalars, workspace, workspace_bytes,
                       VQ_ALGO_TC, stream, shortlist4, reinterpret_cast<long long*>(trace), trace_tiles);
}

int vq_assign_grouped(const float* x, int64_t N, int64_t D, int64_t T, const float* k, int n_vocab, int l_bins,
                      const int64_t* tok, int64_t* q_rel, int64_t* q_abs, float* min_d, double* scalars,
                      void* workspace, size_t workspace_bytes, void* stream_) {
    VQ_REQUIRE(n_vocab > 0 && l_bins > 0 && int64_t(n_vocab) * l_bins < (int64_t(1) << 31), "bad group shape");
    const int K = n_vocab * l_bins;
    if (check_shape(N, D, T, K)) return 1;
    if (N * T == 0) return 0;
    VQ_REQUIRE(x && k && tok && q_rel && q_abs && workspace, "null pointer");
    VQ_REQUIRE(D <= 512, "emb_width > 512 is not supported by the grouped kernel");
    VQ_REQUIRE(vq_device_supported(), "this library only runs on compute capability 10.x (B200); no fallback exists");
    VQ_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    AssignWorkspace w = carve_workspace(workspace, N * T, K, int(D));
    VQ_REQUIRE(workspace_bytes >= w.bytes, "workspace too small (see vq_workspace_bytes)");
    VQ_CUDA_OK(cudaMemsetAsync(w.hdr, 0, sizeof(AssignHeader), stream));
    codebook_prepare_kernel<<<(w.Kp * 32 + 255) / 256, 256, 0, stream>>>(k, K, int(D), w.Kp, w.Dp, w.ee, w.hn, nullptr, w.hdr);
    VQ_CUDA_OK(cudaGetLastError());
    const int grid = int(std::min<int64_t>(16 * int64_t(num_sms()), (N * T + L_WARPS - 1) / L_WARPS));
    const bool vec = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(k) & 15) == 0);
    auto launch = [&](auto kernel) {
        kernel<<<grid, L_WARPS * 32, 0, stream>>>(x, N, int(D), T, k, (const float*)w.ee, n_vocab, l_bins, tok, q_rel, q_abs, min_d, scalars);
    };
    if (vec && D <= 128) launch(assign_grouped_kernel<true, 1>);
    else if (vec) launch(assign_grouped_kernel<true, 4>);
    else if (D <= 128) launch(assign_grouped_kernel<false, 1>);
    else launch(assign_grouped_kernel<false, 4>);
    VQ_CUDA_OK(cudaGetLastError());
    return 0;
}

int vq_gather_st_fwd(const float* x, const int64_t* idx, const float* mask, const float* k, int64_t N, int64_t D,
                     int64_t T, int K, float* x_q, double* scalars, float* results, void* stream) {
    if (check_shape(N, D, T, K)) return 1;
    VQ_REQUIRE(scalars && results, "scalars/results must not be null");
    if (N * T == 0) return 0;
    VQ_REQUIRE(x && idx && k && x_q, "null pointer");
    VQ_REQUIRE(vq_device_supported(), "this library only runs on compute capability 10.x (B200); no fallback exists");
    return launch_gather<GM_FWD>(x, idx, mask, k, nullptr, nullptr, N, int(D), T, K, x_q, scalars, results,
                                 static_cast<cudaStream_t>(stream));
}

int vq_gather_st_fwd_ema_supported(int64_t D, int64_t T, int K) {
    return (K > 0 && K <= FE_KMAX && D > 0 && D <= 512 && T > 0 && T % 4 == 0 && T < (int64_t(1) << 31) &&
            fe_smem_bytes(K) <= size_t(227 * 1024) - 2048) ? 1 : 0;
}

int vq_gather_st_fwd_ema(const float* x, const int64_t* idx, const float* mask, const float* k, int64_t N, int64_t D,
                         int64_t T, int K, float* x_q, double* scalars, float* results, float* stats, void* stream) {
    if (check_shape(N, D, T, K)) return 1;
    VQ_REQUIRE(scalars && results && stats, "scalars/results/stats must not be null");
    if (N * T == 0) return 0;
    VQ_REQUIRE(x && idx && k && x_q, "null pointer");
    VQ_REQUIRE(vq_device_supported(), "this library only runs on compute capability 10.x (B200); no fallback exists");
    VQ_REQUIRE(vq_gather_st_fwd_ema_supported(D, T, K), "fused forward + EMA needs K <= 512, D <= 512, T % 4 == 0");
    VQ_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(x_q) & 15) == 0, "x / x_q must be 16-byte aligned");
    const int64_t units = N * ((T + G_TT - 1) / G_TT);
    VQ_REQUIRE(units < (int64_t(1) << 31), "too many tiles");
    VQ_CUDA_OK(ensure_dynamic_smem(gather_fwd_ema_kernel, int(227 * 1024 - 2048)));
    const int slices = int((D + FE_DS - 1) / FE_DS);
    const int gx = int(std::min<int64_t>(units, std::max(1, num_sms() / slices)));
    gather_fwd_ema_kernel<<<dim3(gx, slices), FE_THREADS, fe_smem_bytes(K), static_cast<cudaStream_t>(stream)>>>(
        x, idx, mask, k, int(N), int(D), int(T), K, x_q, scalars, results, stats, (unsigned int)(gx * slices));
    VQ_CUDA_OK(cudaGetLastError());
    return 0;
}

int vq_gather_st_bwd(const float* x, const int64_t* idx, const float* mask, const float* k, const float* grad_xq,
                     const float* grad_commit, const double* scalars, int64_t N, int64_t D, int64_t T, int K,
                     float* grad_x, void* stream) {
    if (check_shape(N, D, T, K)) return 1;
    if (N * T == 0) return 0;
    VQ_REQUIRE(x && idx && k && grad_x && grad_commit && scalars, "null pointer");
    VQ_REQUIRE(vq_device_supported(), "this library only runs on compute capability 10.x (B200); no fallback exists");
    return launch_gather<GM_BWD>(x, idx, mask, k, grad_xq, grad_commit, N, int(D), T, K, grad_x,
                                 const_cast<double*>(scalars), nullptr, static_cast<cudaStream_t>(stream));
}

int vq_decode(const int64_t* idx, const float* k, int64_t N, int64_t D, int64_t T, int K, float* x_d, void* stream) {
    if (check_shape(N, D, T, K)) return 1;
    if (N * T == 0) return 0;
    VQ_REQUIRE(idx && k && x_d, "null pointer");
    VQ_REQUIRE(vq_device_supported(), "this library only runs on compute capability 10.x (B200); no fallback exists");
    return launch_gather<GM_DECODE>(nullptr, idx, nullptr, k, nullptr, nullptr, N, int(D), T, K, x_d, nullptr, nullptr,
                                    static_cast<cudaStream_t>(stream));
}

int vq_ema_accumulate(const float* x, const int64_t* idx, const float* mask, int64_t N, int64_t D, int64_t T, int K,
                      float* stats, void* scratch, void* stream_) {
    if (check_shape(N, D, T, K)) return 1;
    if (N * T == 0) return 0;
    VQ_REQUIRE(x && idx && stats, "null pointer");
    VQ_REQUIRE(vq_device_supported(), "this library only runs on compute capability 10.x (B200); no fallback exists");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int64_t tiles = N * ((T + E_TT - 1) / E_TT);
    const int64_t tiles_r = N * ((T + ER_TT - 1) / ER_TT);
    unsigned char* flags = nullptr;
    if (mask && scratch) {          // one byte per 64-frame tile: tiles without a valid frame are skipped
        flags = static_cast<unsigned char*>(scratch);
        const int tpu = int((T + ER_TT - 1) / ER_TT);
        ema_tile_flags_kernel<<<unsigned((tiles_r * 32 + 255) / 256), 256, 0, stream>>>(mask, N, T, tpu, flags);
        VQ_CUDA_OK(cudaGetLastError());
    }
    if (K < (1 << 24) && er_smem_bytes<true>(K) <= ER_DYN_SMEM_MAX) {
        // private [K][64] slab per (row chunk, 64-deep slice); one CTA per SM
        VQ_CUDA_OK(ensure_dynamic_smem(ema_accumulate_runs_kernel<true>, ER_DYN_SMEM_MAX));
        const int slices = int((D + ErCfg<true>::DW - 1) / ErCfg<true>::DW);
        const int gx = int(std::min<int64_t>(tiles_r, std::max<int64_t>(1, VQ_K3_OCC * num_sms() / slices)));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(gx, slices);
        cfg.blockDim = dim3(ER_THREADS);
        cfg.dynamicSmemBytes = er_smem_bytes<true>(K);
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = flags ? 1 : 0;        // dependent of the tile-flag kernel: its slab initialisation overlaps that kernel
        VQ_CUDA_OK(cudaLaunchKernelEx(&cfg, ema_accumulate_runs_kernel<true>, x, idx, mask, N, int(D), T, K, stats,
                                      (const unsigned char*)flags));
    } else if (K < (1 << 24) && D <= ErCfg<false>::DW) {
        VQ_CUDA_OK(ensure_dynamic_smem(ema_accumulate_runs_kernel<false>, ER_DYN_SMEM_MAX));
        const int grid = int(std::min<int64_t>(tiles_r, num_sms()));
        ema_accumulate_runs_kernel<false><<<grid, ER_THREADS, er_smem_bytes<false>(K), stream>>>(x, idx, mask, N, int(D), T, K, stats, flags);
    } else {
        int grid = int(std::min<int64_t>(tiles, int64_t(num_sms()) * 8));
        ema_accumulate_global_kernel<<<grid, E_THREADS, 0, stream>>>(x, idx, mask, N, int(D), T, K, stats);
    }
    VQ_CUDA_OK(cudaGetLastError());
    return 0;
}

int vq_ema_finalize(const float* stats, const float* k_rand, const float* k_old, float* k, float* k_sum, float* k_elem, int K, int D,
                    double mu, double threshold, double laplace_eps, double* scalars, float* results,
                    int64_t* used_curr, void* stream_) {
    VQ_REQUIRE(K > 0 && D > 0, "bad codebook shape");
    VQ_REQUIRE(stats && k_rand && k_old && k && k_sum && k_elem && scalars && results, "null pointer");
    VQ_REQUIRE(vq_device_supported(), "this library only runs on compute capability 10.x (B200); no fallback exists");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const float* counts = stats + size_t(K) * D;
    int g0 = std::min(64, (K + 255) / 256);
    ema_count_total_kernel<<<g0, 256, 0, stream>>>(counts, k_elem, float(mu), float(1.0 - mu), K, scalars);
    VQ_CUDA_OK(cudaGetLastError());
    int grid = std::min((K + 7) / 8, num_sms() * 8);
    // mu and (1 - mu) are rounded to FP32 independently, like `mu * t + (1. - mu) * s` on FP32 tensors
    ema_finalize_kernel<<<grid, 256, 0, stream>>>(stats, k_rand, k_old, k, k_sum, k_elem, K, D, float(mu), float(1.0 - mu),
                                                  float(threshold), float(laplace_eps), scalars, results, used_curr,
                                                  (unsigned int)grid);
    VQ_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- EMA statistics exchange over NVLink peer memory (k3_p2p.cuh)
size_t vq_p2p_region_bytes(int k_bins, int emb_width) {
    return (k_bins > 0 && emb_width > 0) ? p2p_region_bytes(k_bins, emb_width) : 0;
}

int vq_p2p_alloc(size_t bytes, void** region_out, unsigned char* handle64_out) {
    VQ_REQUIRE(region_out && handle64_out && bytes >= P2P_FLAG_BYTES, "bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    void* ptr = nullptr;
    VQ_CUDA_OK(cudaMalloc(&ptr, bytes));
    VQ_CUDA_OK(cudaMemset(ptr, 0, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
    if (e != cudaSuccess) { cudaFree(ptr); return fail("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e)); }
    memcpy(handle64_out, &h, 64);
    *region_out = ptr;
    return 0;
}

int vq_p2p_open(const unsigned char* handle64, void** region_out) {
    VQ_REQUIRE(handle64 && region_out, "null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    VQ_CUDA_OK(cudaIpcOpenMemHandle(region_out, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int vq_p2p_close(void* peer_region) {
    if (peer_region) VQ_CUDA_OK(cudaIpcCloseMemHandle(peer_region));
    return 0;
}

int vq_p2p_free(void* region) {
    if (region) VQ_CUDA_OK(cudaFree(region));
    return 0;
}

float* vq_p2p_stats_slot(void* region, unsigned int step, int k_bins, int emb_width) {
    return region ? p2p_stats_slot(region, step, p2p_stats_floats(k_bins, emb_width)) : nullptr;
}

float* vq_p2p_krand_slot(void* region, unsigned int step, int k_bins, int emb_width) {
    return region ? p2p_krand_slot(region, step, p2p_stats_floats(k_bins, emb_width), size_t(k_bins) * emb_width) : nullptr;
}

static int p2p_peers(void* const* regions, int n_ranks, int rank, P2PPeers& peers) {
    VQ_REQUIRE(regions, "null pointer");
    VQ_REQUIRE(n_ranks >= 1 && n_ranks <= P2P_MAX_RANKS && rank >= 0 && rank < n_ranks, "bad rank / world size (at most 16 ranks)");
    VQ_REQUIRE(vq_device_supported(), "this library only runs on compute capability 10.x (B200); no fallback exists");
    for (int r = 0; r < P2P_MAX_RANKS; ++r) peers.region[r] = r < n_ranks ? regions[r] : nullptr;
    for (int r = 0; r < n_ranks; ++r) VQ_REQUIRE(peers.region[r] != nullptr, "a peer region is not mapped");
    return 0;
}

int vq_p2p_publish(void* const* regions, int n_ranks, int rank, unsigned int step, void* stream) {
    P2PPeers peers;
    if (p2p_peers(regions, n_ranks, rank, peers)) return 1;
    p2p_publish_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(peers, n_ranks, rank, step);
    VQ_CUDA_OK(cudaGetLastError());
    return 0;
}

int vq_p2p_collect(void* const* regions, int n_ranks, int rank, unsigned int step, int k_bins, int emb_width,
                   float* stats_out, float* k_rand_out, void* stream_) {
    VQ_REQUIRE(stats_out && k_rand_out, "null pointer");
    P2PPeers peers;
    if (p2p_peers(regions, n_ranks, rank, peers)) return 1;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t sf = p2p_stats_floats(k_bins, emb_width), kf = size_t(k_bins) * emb_width;
    p2p_wait_kernel<<<1, 32, 0, stream>>>(static_cast<const unsigned*>(regions[rank]), n_ranks, step);
    const int grid = int(std::min<size_t>((sf / 4 + 255) / 256, size_t(num_sms()) * 4));
    p2p_reduce_kernel<<<std::max(grid, 1), 256, 0, stream>>>(peers, n_ranks, step, sf, kf, stats_out, k_rand_out);
    VQ_CUDA_OK(cudaGetLastError());
    return 0;
}

int vq_p2p_exchange(void* const* regions, int n_ranks, int rank, unsigned int step, int k_bins, int emb_width,
                    float* stats_out, float* k_rand_out, void* stream) {
    if (vq_p2p_publish(regions, n_ranks, rank, step, stream)) return 1;
    return vq_p2p_collect(regions, n_ranks, rank, step, k_bins, emb_width, stats_out, k_rand_out, stream);
}

int vq_restart_rows_device(const float* x, const float* mask, int64_t N, int64_t D, int64_t T, int K, uint64_t seed,
                           const uint64_t* seed_dev, float* out, void* scratch, void* stream_) {
    if (check_shape(N, D, T, K)) return 1;
    VQ_REQUIRE(x && out && scratch, "null pointer");
    VQ_REQUIRE(vq_device_supported(), "this library only runs on compute capability 10.x (B200); no fallback exists");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    long long* prefix = static_cast<long long*>(scratch);
    if (N == 0 || T == 0) {
        VQ_CUDA_OK(cudaMemsetAsync(out, 0, size_t(K) * D * sizeof(float), stream));
        return 0;
    }
    restart_count_kernel<<<unsigned((N * 32 + 255) / 256), 256, 0, stream>>>(mask, N, T, prefix);
    VQ_CUDA_OK(cudaGetLastError());
    restart_scan_kernel<<<1, 1024, 0, stream>>>(prefix, N);
    VQ_CUDA_OK(cudaGetLastError());
    restart_select_kernel<<<K, 128, 0, stream>>>(x, mask, prefix, N, int(D), T, K, seed, seed_dev, out);
    VQ_CUDA_OK(cudaGetLastError());
    return 0;
}

int vq_gather_rows(const float* x, const int64_t* rows, int64_t n_rows, int64_t N, int64_t D, int64_t T, float* out,
                   void* stream) {
    if (n_rows == 0) return 0;
    VQ_REQUIRE(x && rows && out && n_rows > 0 && n_rows < (int64_t(1) << 31), "bad arguments");
    VQ_REQUIRE(vq_device_supported(), "this library only runs on compute capability 10.x (B200); no fallback exists");
    gather_rows_kernel<<<unsigned(n_rows), 128, 0, static_cast<cudaStream_t>(stream)>>>(x, rows, n_rows, N, int(D), T, out);
    VQ_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Host-buffer path: H2D copy -> K1 -> D2H copy, chunked by utterance and double-buffered on two streams.
struct vq_host_ctx {
    int device = 0;
    int64_t max_rows = 0;
    int K = 0, D = 0;
    float* d_x[2] = {nullptr, nullptr};
    int64_t* d_idx[2] = {nullptr, nullptr};
    void* d_ws[2] = {nullptr, nullptr};
    double* d_scalars[2] = {nullptr, nullptr};
    size_t ws_bytes = 0;
    int64_t chunk_rows = 0;
    float* d_k = nullptr;
    float* h_x = nullptr;
    int64_t* h_idx = nullptr;
    double* h_scalars = nullptr;
    cudaStream_t streams[2] = {nullptr, nullptr};
    bool prepared[2] = {false, false};      // workspace b already holds the current codebook's operands
    // compact-code path (vq_encode_host_u16)
    unsigned short* d_codes[2] = {nullptr, nullptr};
    int64_t* d_off[2] = {nullptr, nullptr};
    int64_t* h_off[2] = {nullptr, nullptr};
    int64_t off_cap = 0;                    // utterances per chunk the offset buffers hold
    unsigned short* h_codes = nullptr;
};

void vq_host_ctx_destroy(vq_host_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (int i = 0; i < 2; ++i) {
        if (c->streams[i]) cudaStreamSynchronize(c->streams[i]);
        cudaFree(c->d_x[i]); cudaFree(c->d_idx[i]); cudaFree(c->d_ws[i]); cudaFree(c->d_scalars[i]);
        if (c->streams[i]) cudaStreamDestroy(c->streams[i]);
    }
    for (int i = 0; i < 2; ++i) { cudaFree(c->d_codes[i]); cudaFree(c->d_off[i]); cudaFreeHost(c->h_off[i]); }
    cudaFree(c->d_k);
    cudaFreeHost(c->h_x); cudaFreeHost(c->h_idx); cudaFreeHost(c->h_scalars); cudaFreeHost(c->h_codes);
    delete c;
}

vq_host_ctx* vq_host_ctx_create(int device, int64_t max_rows, int K, int D) {
    if (max_rows <= 0 || K <= 0 || D <= 0) { fail("vq_host_ctx_create: bad arguments%s"); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { fail("vq_host_ctx_create: cudaSetDevice failed%s"); return nullptr; }
    if (!vq_device_supported()) { fail("vq_host_ctx_create: needs compute capability 10.x (B200); no fallback exists%s"); return nullptr; }
    vq_host_ctx* c = new (std::nothrow) vq_host_ctx();
    if (!c) return nullptr;
    c->device = device; c->max_rows = max_rows; c->K = K; c->D = D;
    // each in-flight chunk holds at most half of the job (rounded up), so two buffers cover any split
    c->chunk_rows = max_rows;
    c->ws_bytes = vq_workspace_bytes(1, c->chunk_rows, K, D);
    bool ok = true;
    for (int i = 0; i < 2 && ok; ++i) {
        ok = ok && cudaMalloc(&c->d_x[i], size_t(c->chunk_rows) * D * 4) == cudaSuccess;
        ok = ok && cudaMalloc(&c->d_idx[i], size_t(c->chunk_rows) * 8) == cudaSuccess;
        ok = ok && cudaMalloc(&c->d_ws[i], c->ws_bytes) == cudaSuccess;
        ok = ok && cudaMalloc(&c->d_scalars[i], VQ_NUM_SCALARS * 8) == cudaSuccess;
        ok = ok && cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking) == cudaSuccess;
    }
    ok = ok && cudaMalloc(&c->d_k, size_t(K) * D * 4) == cudaSuccess;
    ok = ok && cudaMallocHost(&c->h_x, size_t(max_rows) * D * 4) == cudaSuccess;
    ok = ok && cudaMallocHost(&c->h_idx, size_t(max_rows) * 8) == cudaSuccess;
    ok = ok && cudaMallocHost(&c->h_scalars, 2 * VQ_NUM_SCALARS * 8) == cudaSuccess;
    if (!ok) {
        fail("vq_host_ctx_create: allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        vq_host_ctx_destroy(c);
        return nullptr;
    }
    return c;
}

uint16_t* vq_host_ctx_codes_staging(vq_host_ctx* c) {
    if (!c) return nullptr;
    if (!c->h_codes && cudaMallocHost(&c->h_codes, size_t(c->max_rows) * 2) != cudaSuccess) { fail("vq_host_ctx_codes_staging: allocation failed%s"); return nullptr; }
    return c->h_codes;
}
float* vq_host_ctx_x_staging(vq_host_ctx* c) { return c ? c->h_x : nullptr; }
int64_t* vq_host_ctx_idx_staging(vq_host_ctx* c) { return c ? c->h_idx : nullptr; }

int vq_host_ctx_set_codebook(vq_host_ctx* c, const float* k_host) {
    VQ_REQUIRE(c && k_host, "null pointer");
    VQ_CUDA_OK(cudaSetDevice(c->device));
    VQ_CUDA_OK(cudaMemcpy(c->d_k, k_host, size_t(c->K) * c->D * 4, cudaMemcpyHostToDevice));
    c->prepared[0] = c->prepared[1] = false;
    return 0;
}

int vq_encode_host(vq_host_ctx* c, const float* x_host, int64_t N, int64_t T, int64_t* idx_host, double* sum_min_d_host) {
    VQ_REQUIRE(c && x_host && idx_host, "null pointer");
    VQ_REQUIRE(N >= 0 && T >= 0 && N * T <= c->max_rows, "job larger than the context was created for");
    VQ_CUDA_OK(cudaSetDevice(c->device));
    if (N * T == 0) { if (sum_min_d_host) *sum_min_d_host = 0.0; return 0; }
    // split by utterance into pieces of ~32 MB of latents so copy-in, K1 and copy-out of neighbouring pieces overlap
    const int64_t bytes_per_utt = int64_t(c->D) * T * 4;
    int64_t utt_per_chunk = std::max<int64_t>(1, (int64_t(32) << 20) / std::max<int64_t>(1, bytes_per_utt));
    utt_per_chunk = std::min(utt_per_chunk, N);
    const int64_t n_chunks = (N + utt_per_chunk - 1) / utt_per_chunk;
    // device buffers hold one chunk each; chunk offsets inside the buffer rotate so two chunks are in flight
    double total = 0.0;
    cudaEvent_t done[2];
    VQ_CUDA_OK(cudaEventCreateWithFlags(&done[0], cudaEventDisableTiming));
    VQ_CUDA_OK(cudaEventCreateWithFlags(&done[1], cudaEventDisableTiming));
    int rc = 0;
    const bool want_sum = sum_min_d_host != nullptr;     // without it the cheaper non-rescoring kernel variant runs
    auto cuda_ok = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && !rc) {
            snprintf(err_buf(), 512, "vq_encode_host: %s failed: %s", what, cudaGetErrorString(e));
            rc = 1;
        }
        return e == cudaSuccess;
    };
    for (int64_t ci = 0; ci < n_chunks && !rc; ++ci) {
        const int b = int(ci & 1);
        cudaStream_t s = c->streams[b];
        const int64_t n0 = ci * utt_per_chunk, nn = std::min(utt_per_chunk, N - n0);
        if (ci >= 2) {   // the buffer's previous job (scalars read-back included) must be finished
            if (!cuda_ok(cudaEventSynchronize(done[b]), "cudaEventSynchronize")) break;
            if (want_sum) total += c->h_scalars[b * VQ_NUM_SCALARS + VQ_S_SUM_MIN_D];
        }
        if (!cuda_ok(cudaMemcpyAsync(c->d_x[b], x_host + n0 * int64_t(c->D) * T, size_t(nn) * bytes_per_utt, cudaMemcpyHostToDevice, s),
                     "cudaMemcpyAsync(x)")) break;
        if (want_sum && !cuda_ok(cudaMemsetAsync(c->d_scalars[b], 0, VQ_NUM_SCALARS * 8, s), "cudaMemsetAsync(scalars)")) break;
        if (vq_assign(c->d_x[b], nn, c->D, T, c->d_k, c->K, c->d_idx[b], nullptr, want_sum ? c->d_scalars[b] : nullptr, c->d_ws[b],
                      c->ws_bytes, VQ_ALGO_AUTO | (c->prepared[b] ? VQ_ALGO_PREPARED : 0), s)) { rc = 1; break; }
        c->prepared[b] = true;
        if (!cuda_ok(cudaMemcpyAsync(idx_host + n0 * T, c->d_idx[b], size_t(nn) * T * 8, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync(idx)")) break;
        if (want_sum && !cuda_ok(cudaMemcpyAsync(c->h_scalars + b * VQ_NUM_SCALARS, c->d_scalars[b], VQ_NUM_SCALARS * 8,
                                                 cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync(scalars)")) break;
        if (!cuda_ok(cudaEventRecord(done[b], s), "cudaEventRecord")) break;
    }
    for (int b = 0; b < 2; ++b) {
        cudaError_t e = cudaStreamSynchronize(c->streams[b]);
        if (e != cudaSuccess && !rc) rc = fail("vq_encode_host: %s", cudaGetErrorString(e));
    }
    if (!rc && want_sum) {
        const int64_t tail = std::min<int64_t>(2, n_chunks);
        for (int64_t j = n_chunks - tail; j < n_chunks; ++j) total += c->h_scalars[(j & 1) * VQ_NUM_SCALARS + VQ_S_SUM_MIN_D];
    }
    cudaEventDestroy(done[0]); cudaEventDestroy(done[1]);
    if (sum_min_d_host) *sum_min_d_host = total;
    return rc;
}

int vq_encode_host_u16(vq_host_ctx* c, const float* x_host, int64_t N, int64_t T, const int32_t* lengths_host,
                       uint16_t* codes_host, int64_t* total_codes_out) {
    VQ_REQUIRE(c && x_host && codes_host, "null pointer");
    VQ_REQUIRE(N >= 0 && T >= 0 && N * T <= c->max_rows, "job larger than the context was created for");
    VQ_REQUIRE(c->K <= 65536, "codes do not fit 16 bits (k_bins > 65536)");
    VQ_CUDA_OK(cudaSetDevice(c->device));
    if (total_codes_out) *total_codes_out = 0;
    if (N * T == 0) return 0;
    const int64_t bytes_per_utt = int64_t(c->D) * T * 4;
    int64_t utt_per_chunk = std::max<int64_t>(1, (int64_t(32) << 20) / std::max<int64_t>(1, bytes_per_utt));
    utt_per_chunk = std::min(utt_per_chunk, N);
    const int64_t n_chunks = (N + utt_per_chunk - 1) / utt_per_chunk;
    for (int i = 0; i < 2; ++i)
        if (!c->d_codes[i]) VQ_CUDA_OK(cudaMalloc(&c->d_codes[i], size_t(c->chunk_rows) * 2));
    if (c->off_cap < utt_per_chunk + 1) {
        for (int i = 0; i < 2; ++i) {
            VQ_CUDA_OK(cudaStreamSynchronize(c->streams[i]));
            cudaFree(c->d_off[i]); cudaFreeHost(c->h_off[i]);
            c->d_off[i] = nullptr; c->h_off[i] = nullptr;
            VQ_CUDA_OK(cudaMalloc(&c->d_off[i], size_t(utt_per_chunk + 1) * 8));
            VQ_CUDA_OK(cudaMallocHost(&c->h_off[i], size_t(utt_per_chunk + 1) * 8));
        }
        c->off_cap = utt_per_chunk + 1;
    }
    cudaEvent_t done[2];
    VQ_CUDA_OK(cudaEventCreateWithFlags(&done[0], cudaEventDisableTiming));
    VQ_CUDA_OK(cudaEventCreateWithFlags(&done[1], cudaEventDisableTiming));
    int rc = 0;
    auto cuda_ok = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && !rc) {
            snprintf(err_buf(), 512, "vq_encode_host_u16: %s failed: %s", what, cudaGetErrorString(e));
            rc = 1;
        }
        return e == cudaSuccess;
    };
    int64_t written = 0;                      // codes of the chunks issued so far
    for (int64_t ci = 0; ci < n_chunks && !rc; ++ci) {
        const int b = int(ci & 1);
        cudaStream_t s = c->streams[b];
        const int64_t n0 = ci * utt_per_chunk, nn = std::min(utt_per_chunk, N - n0);
        if (ci >= 2 && !cuda_ok(cudaEventSynchronize(done[b]), "cudaEventSynchronize")) break;   // buffer b (and its pinned offsets) free again
        int64_t* off = c->h_off[b];
        off[0] = 0;
        for (int64_t i = 0; i < nn; ++i) {
            int64_t len = lengths_host ? int64_t(lengths_host[n0 + i]) : T;
            if (len < 0 || len > T) { rc = fail("vq_encode_host_u16: a length is outside [0, T]%s"); break; }
            off[i + 1] = off[i] + len;
        }
        if (rc) break;
        const int64_t chunk_codes = off[nn];
        if (!cuda_ok(cudaMemcpyAsync(c->d_x[b], x_host + n0 * int64_t(c->D) * T, size_t(nn) * bytes_per_utt, cudaMemcpyHostToDevice, s), "cudaMemcpyAsync(x)")) break;
        if (!cuda_ok(cudaMemcpyAsync(c->d_off[b], off, size_t(nn + 1) * 8, cudaMemcpyHostToDevice, s), "cudaMemcpyAsync(offsets)")) break;
        if (vq_assign(c->d_x[b], nn, c->D, T, c->d_k, c->K, c->d_idx[b], nullptr, nullptr, c->d_ws[b], c->ws_bytes,
                      VQ_ALGO_AUTO | (c->prepared[b] ? VQ_ALGO_PREPARED : 0), s)) { rc = 1; break; }
        c->prepared[b] = true;
        if (chunk_codes > 0) {
            const dim3 grid(unsigned(std::min<int64_t>((T + 255) / 256, 64)), unsigned(nn));
            pack_codes_u16_kernel<<<grid, 256, 0, s>>>(c->d_idx[b], T, c->d_off[b], c->d_codes[b]);
            if (!cuda_ok(cudaGetLastError(), "pack_codes_u16_kernel")) break;
            if (!cuda_ok(cudaMemcpyAsync(codes_host + written, c->d_codes[b], size_t(chunk_codes) * 2, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync(codes)")) break;
        }
        written += chunk_codes;
        if (!cuda_ok(cudaEventRecord(done[b], s), "cudaEventRecord")) break;
    }
    for (int b = 0; b < 2; ++b) {
        cudaError_t e = cudaStreamSynchronize(c->streams[b]);
        if (e != cudaSuccess && !rc) rc = fail("vq_encode_host_u16: %s", cudaGetErrorString(e));
    }
    cudaEventDestroy(done[0]); cudaEventDestroy(done[1]);
    if (total_codes_out) *total_codes_out = written;
    return rc;
}

}  // extern "C"
