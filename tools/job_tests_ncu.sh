mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02b_gputests.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/r02b_gputests.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b_launches_frame.csv python profiles/frame_profile.py 3 > gpurun_out/fp.log 2>&1; tail -2 gpurun_out/fp.log
ncu --set full --clock-control none --import-source on -k regex:k_ -s 26 -c 13 -f -o gpurun_out/r02b_frame python profiles/frame_profile.py 3 > gpurun_out/fp2.log 2>&1; tail -3 gpurun_out/fp2.log
ls -la gpurun_out
