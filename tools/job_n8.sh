# GPU job (8 GPUs): exchange micro-benchmark for two block counts, then the bench line at N=8.  usage: bash tools/job_n8.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
for b in 64 128; do GSR_AR_BLOCKS=$b python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/probe_symm.py 2>&1 | grep "library\|nccl all"; done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --no-extras > gpurun_out/${tag}_bench_n8.json 2> gpurun_out/${tag}_bench_n8.err
echo "bench exit $?"; tail -2 gpurun_out/${tag}_bench_n8.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_bench_n8.json').read().strip().splitlines()[-1])
print('N=8 value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'kernel_ms',d.get('kernel_ms'),'comm_ms',d.get('comm_ms'),'e2e',round(d.get('e2e',{}).get('value',0),1),'exchange',d['run'].get('exchange'))
PY
