# GPU job: full GPU test suite, then the default bench line.  usage: bash tools/job_tests_bench.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_gputests.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/${tag}_gputests.log
timeout 600 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench exit $?"; tail -3 gpurun_out/${tag}_bench_n1.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_bench_n1.json').read().strip().splitlines()[-1])
print('value',d.get('value'),'ms/step',d.get('ms_per_step'))
print('stages',{k:v['ms'] for k,v in d.get('stages',{}).items()})
print('e2e',d.get('e2e',{}).get('value'),'m1',d.get('m1',{}).get('m1_single_call_fps'),'m2',d.get('m2',{}).get('two_pass_fps'))
print('parity',d.get('parity_check',{}).get('ok'))
PY
