# GPU job: `ncu --set full` capture of every library kernel of ONE sequential C-main frame (the third of three), plus the
# launch list of the same workload.  usage: bash tools/job_ncu_full.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_frame.csv python profiles/frame_profile.py 3 > gpurun_out/fp.log 2>&1; tail -1 gpurun_out/fp.log
# library kernels per frame: preprocess_fwd, 4 sort passes, partition, render_fwd, render_bwd, preprocess_bwd = 9
ncu --set full --clock-control none --import-source on -k regex:k_ -s 18 -c 9 -f -o gpurun_out/${tag}_frame python profiles/frame_profile.py 3 > gpurun_out/fp2.log 2>&1; tail -2 gpurun_out/fp2.log
ncu -i gpurun_out/${tag}_frame.ncu-rep --page raw --csv > gpurun_out/${tag}_frame_raw.csv 2>/dev/null
ls -la gpurun_out | grep ${tag}
