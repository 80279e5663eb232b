# GPU job (N GPUs): the closing measurements at N ranks — C-main bench line (with e2e) and the C5 line (no e2e leg).
# At N = 2 the multi-rank parity test runs first.  usage: bash tools/job_final_multi.sh <tag> <N>
tag=${1:-x}; N=${2:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_multirank.py -q -x 2>&1 | tail -3; fi
run() {  # name, args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --no-extras $2 > gpurun_out/${tag}_bench_$1_n${N}.json 2> gpurun_out/${tag}_bench_$1_n${N}.err
  echo "bench $1 exit $?"; grep -v "OMP_NUM\|^\*\*\*\|bench +\|^$" gpurun_out/${tag}_bench_$1_n${N}.err | tail -3 | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_bench_$1_n${N}.json').read().strip().splitlines()[-1])
    print('$1 N=$N value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'kernel_ms',d.get('kernel_ms'),'comm_ms',d.get('comm_ms'),'e2e',round(d.get('e2e',{}).get('value',0),1),'exchange',d['run'].get('exchange'), d['run'].get('exchange_fallback_reason'))
except Exception as e: print('no line', e)
PY
}
run cmain ""
run c5 "--config c5 --steps 3 --repeats 3 --no-e2e"
