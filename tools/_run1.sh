set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_sweepx.py tests/test_gpu_edge_cases.py tests/test_gpu_examples.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r1b_tests_sym.log
for v in 0 2 3 4; do
  MB_STATIC_SYM=$v timeout 300 python bench.py --no-e2e --cpu-sample 200 --steps 10 --warmup 3 > gpurun_out/r1b_ab_static_$v.json 2> gpurun_out/r1b_ab_static_$v.err
done
cat gpurun_out/r1b_tests_sym.log
python - <<'PY'
import json
for v in (0,2,3,4):
    try:
        d=json.loads(open('gpurun_out/r1b_ab_static_%d.json'%v).read().strip().splitlines()[-1])
        print(v, d['value'], d['breakdown_ms'], d['clocks'])
    except Exception as e: print(v,'ERR',e)
PY
