O=gpurun_out
python scripts/ncu_step.py 2048 > $O/step2048_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_step2048.csv python scripts/ncu_step.py 2048 > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"bc_(forward|backward)(_look)?_kernel" -c 40 -f -o $O/r02_step python scripts/ncu_step.py 2048 > $O/ncu_full.log 2>&1
tail -n 2 $O/step2048_plain.log
