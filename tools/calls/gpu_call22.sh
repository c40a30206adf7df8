#!/bin/bash
# round 2, call 22 (the last 3 GPU-minutes): whole steps as CUDA graphs on latency-bound grids (step_graphed) on hardware -- the
# named configs at native size, direct against graphed (bit-exact both ways, ms per step of each), then as much of the suite as fits
out=gpurun_out; tag=r2c22; mkdir -p $out
timeout 165 python -m pytest tests/test_gpu_native_configs.py tests/test_gpu_parity.py tests/test_gpu_outputs.py tests/test_gpu_zz_more_cases.py tests/test_gpu_dropin.py \
  -m gpu -v -rxXs -p no:cacheprovider --deselect tests/test_gpu_native_configs.py::test_unstable_jet_long_run_matches_oracle_and_conserves_volume > $out/${tag}_pytest.log 2>&1
echo "pytest: exit $?" >> $out/${tag}_pytest.log
grep -c PASSED $out/${tag}_pytest.log; grep -E "FAILED|ERROR|exit" $out/${tag}_pytest.log | head -20
cp gpurun_out/native_configs.json $out/${tag}_native_configs.json 2>/dev/null
python - <<'P'
import json
try:
    d = json.load(open("gpurun_out/native_configs.json"))
    for k, v in sorted(d.items()):
        if "ms_per_step" in v:
            print(k, "graphed %.4f ms" % v["ms_per_step"], "direct %.4f ms" % v.get("ms_per_step_direct", -1), "graph steps", v.get("graph_steps"))
except Exception as e:
    print("no native_configs.json:", e)
P
