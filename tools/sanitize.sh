#!/bin/bash
# compute-sanitizer passes over small cases of every kernel (memcheck: out-of-bounds / misaligned accesses; racecheck:
# shared-memory hazards in the fused kernel, the projection and the forward FFTs; synccheck: barrier misuse).
# usage (on the GPU box): tools/sanitize.sh > gpurun_out/sanitize.log 2>&1
set -u
cat > /tmp/san_case.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from bioem_b200 import api
from bioem_b200.cases import build_case
for name in sys.argv[1:]:
    cd = build_case(name)
    hi, parts = api.inputs_for_case(cd)
    e = api.Engine(hi.cfg)
    e.upload_all(hi, parts)
    e.run()
    pm, pa = e.download()
    if hi.cfg.writeAngles:
        e.download_top_angles(3)
    print(name, "ok", float(pm["Constoadd"][0]))
    e.close()
PY
for tool in memcheck racecheck synccheck; do
  echo "=== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san_case.py toy32 toy32pts toy36g2 2>&1 | tail -6
  echo "rc=$?"
done
echo "=== memcheck, N = 224 (cfg2_slice, one orientation)"
cat > /tmp/san_224.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from bioem_b200 import api
from bioem_b200.cases import build_case
cd = build_case("cfg2_slice", n_particles=2, n_orient=1)
hi, parts = api.inputs_for_case(cd)
e = api.Engine(hi.cfg); e.upload_all(hi, parts); e.run(); pm, _ = e.download(); print("cfg2_slice ok", float(pm["Constoadd"][0])); e.close()
PY
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san_224.py 2>&1 | tail -4
echo "rc=$?"
echo "=== racecheck, N = 224"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san_224.py 2>&1 | tail -4
echo "rc=$?"
