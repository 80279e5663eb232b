# GPU job: one bench line for a config at N GPUs.  usage: bash tools/job_bench_cfg.sh <tag> <N> <config> [extra bench args]
tag=${1:-x}; N=${2:-1}; cfg=${3:-cmain}; shift 3
mkdir -p gpurun_out
out=gpurun_out/${tag}_bench_${cfg}_n${N}
if [ "$N" = "1" ]; then
  timeout 1200 python bench.py --config $cfg "$@" > $out.json 2> $out.err
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --config $cfg "$@" > $out.json 2> $out.err
fi
echo "bench exit $?"; tail -2 $out.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('$out.json').read().strip().splitlines()[-1])
print('$cfg N=$N value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'kernel_ms',d.get('kernel_ms'),'comm_ms',d.get('comm_ms'),'e2e',round(d.get('e2e',{}).get('value',0),1),'exchange',d['run'].get('exchange'))
print('stages',{k:v['ms'] for k,v in d.get('stages',{}).items()}, 'R_mean', d['run'].get('R_mean'))
PY
