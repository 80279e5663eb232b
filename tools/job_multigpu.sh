# GPU job (N GPUs): multi-rank parity test, then bench at N with NCCL and with the peer exchange.  usage: bash tools/job_multigpu.sh <tag> <N>
tag=${1:-x}; N=${2:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multirank.py -q -x 2>&1 | tail -15
for ex in nccl auto; do
  GSR_EXCHANGE=$ex timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --no-extras > gpurun_out/${tag}_bench_n${N}_${ex}.json 2> gpurun_out/${tag}_bench_n${N}_${ex}.err
  echo "bench $ex exit $?"; tail -2 gpurun_out/${tag}_bench_n${N}_${ex}.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_bench_n${N}_${ex}.json').read().strip().splitlines()[-1])
    print('$ex N=$N value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'kernel_ms',d.get('kernel_ms'),'comm_ms',d.get('comm_ms'),'e2e',round(d.get('e2e',{}).get('value',0),1),'exchange',d['run'].get('exchange'), d['run'].get('exchange_fallback_reason'))
except Exception as e: print('no line', e)
PY
done
