# GPU job (1 GPU): the closing single-GPU measurements — full GPU suite, C-main bench line, the reference arm on the
# same box, the C5 line.  usage: bash tools/job_final_n1.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${tag}_gputests.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${tag}_gputests.log
timeout 600 python bench.py --impl reference > gpurun_out/${tag}_bench_reference_n1.json 2> gpurun_out/${tag}_bench_reference_n1.err; echo "reference exit $?"
timeout 600 python bench.py > gpurun_out/${tag}_bench_cmain_n1.json 2> gpurun_out/${tag}_bench_cmain_n1.err; echo "bench exit $?"
timeout 900 python bench.py --config c5 --steps 3 --repeats 3 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/${tag}_bench_c5_n1.json 2> gpurun_out/${tag}_bench_c5_n1.err; echo "c5 exit $?"
python - <<PY
import json
for f in ('reference','cmain','c5'):
    try:
        d=json.loads(open('gpurun_out/${tag}_bench_%s_n1.json' % f).read().strip().splitlines()[-1])
        print(f,'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'e2e',round(d.get('e2e',{}).get('value',0),1),'m1',d.get('m1',{}).get('m1_single_call_fps'),'m2',d.get('m2',{}).get('two_pass_fps'), 'm3', d.get('m3',{}).get('iterations_per_s'))
        if 'stages' in d: print('   stages',{k:v['ms'] for k,v in d['stages'].items()})
    except Exception as e: print(f,'no line',e)
PY
