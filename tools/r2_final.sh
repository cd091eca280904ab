#!/bin/bash
# final check: what the driver runs at round end -- GPU tests, smoke(), the default bench line
mkdir -p gpurun_out
s=$(date +%s)
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/r02_final_pytest_gpu.log 2>&1
rc=$?; echo "== pytest -m gpu exit $rc ($(( $(date +%s) - s )) s)"; tail -3 gpurun_out/r02_final_pytest_gpu.log
[ $rc -ne 0 ] && grep -E "^E |Error|b2:|FAILED" gpurun_out/r02_final_pytest_gpu.log | head -30
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "bench rc $?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/r02_final_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "model_frac_of_peak", "gpu_launches")}, d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["traffic"], d["e2e"]["value"])
print("decode", {k: d["decode"][k] for k in ("value", "ms_per_token_step")}, d["decode"]["roofline"]["frac"])
print("precise", d["precise"]["value"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], "clocks", d["clocks"])
P
grep -h "\[parity\]" gpurun_out/r02_final_pytest_gpu.log | head -3
exit $rc
