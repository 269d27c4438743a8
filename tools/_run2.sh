set -x
cd $GRAFT_REPO_ROOT
MB_E2E_MIN_NNZ=100 timeout 600 python -m pytest tests/test_gpu_sweepx.py tests/test_gpu_edge_cases.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r1b_tests_ov.log
for c in 0:8 1:8 1:16 1:32; do
  ov=${c%%:*}; ch=${c##*:}
  MB_DEV_OVERLAP=$ov MB_E2E_CHUNKS=$ch timeout 300 python bench.py --no-e2e --cpu-sample 200 --steps 10 --warmup 3 > gpurun_out/r1b_ov_${ov}_${ch}.json 2> gpurun_out/r1b_ov_${ov}_${ch}.err
done
cat gpurun_out/r1b_tests_ov.log
python - <<'PY'
import json
for v in ("0_8","1_8","1_16","1_32"):
    try:
        d=json.loads(open('gpurun_out/r1b_ov_%s.json'%v).read().strip().splitlines()[-1])
        print(v, d['value'], d['ms_per_step'], d['breakdown_ms']['element_kernels'], d['breakdown_ms']['segmented_reduction'], d['clocks'], d['gpu_launches'])
    except Exception as e: print(v,'ERR',e)
PY
tail -3 gpurun_out/r1b_ov_1_8.err
