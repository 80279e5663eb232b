# GPU job (N GPUs): multi-rank parity test, then bench at N (default config) and with K = N keyframes (one per rank).
# usage: bash tools/job_multigpu2.sh <tag> <N>
tag=${1:-x}; N=${2:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multirank.py -q -x 2>&1 | tail -15
run() {  # name, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --no-extras $2 > gpurun_out/${tag}_bench_n${N}_$1.json 2> gpurun_out/${tag}_bench_n${N}_$1.err
  echo "bench $1 exit $?"; grep -v "OMP_NUM\|^\*\*\*\|bench +" gpurun_out/${tag}_bench_n${N}_$1.err | tail -3 | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_bench_n${N}_$1.json').read().strip().splitlines()[-1])
    print('$1 N=$N value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'kernel_ms',d.get('kernel_ms'),'comm_ms',d.get('comm_ms'),'e2e',round(d.get('e2e',{}).get('value',0),1),'exchange',d['run'].get('exchange'), d['run'].get('exchange_fallback_reason'))
except Exception as e: print('no line', e)
PY
}
run default ""
if [ "$N" != "8" ]; then run k$N "--keyframes $N --no-e2e"; fi
