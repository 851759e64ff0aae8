#!/bin/bash
# in-situ duration of one MMA batch (issuer-side probe, event 15 - event 10) under the timing experiments of experiment.sh
for e in ${EXPERIMENTS:-0 1 2 8 32 43}; do
  VQ_EXPERIMENT=$e python speech-masters-thesis_b200/build.py --force > /dev/null 2>&1
  VQ_EXPERIMENT=$e python tools/tc_timeline.py > /tmp/tl.log 2>&1
  python - "$e" <<'PY'
import json, sys
d = json.load(open("gpurun_out/tc_timeline.json"))["fine"]
for name, sel in (("nt0", 0), ("nt2", 2)):
    v = [r[5] - r[0] for i, r in enumerate(d) if i % 4 == sel and r[5] > 0 and r[0] > 0]
    per = [d[i + 4][0] - d[i][0] for i in range(sel, len(d) - 4, 4) if d[i][0] > 0 and d[i + 4][0] > 0]
    print(f"experiment {sys.argv[1]} {name}: MMA batch go->done (issuer) {sum(v)/max(len(v),1):7.0f} cycles, tile period {sum(per)/max(len(per),1):7.0f}")
PY
done
python speech-masters-thesis_b200/build.py --force > /dev/null 2>&1
