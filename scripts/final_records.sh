# Evidence of the final build, one GPU.  bash scripts/final_records.sh   (everything lands in gpurun_out/)
set -x
O=gpurun_out
python scripts/ncu_step.py 2048 > $O/step2048_plain.log 2>&1 || exit 1      # exits 0 without ncu first
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_step2048.csv \
    python scripts/ncu_step.py 2048 > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"bc_(forward|backward)(_look)?_kernel" -c 40 -f \
    -o $O/r02_step python scripts/ncu_step.py 2048 > $O/ncu_full.log 2>&1
python scripts/probe_lines.py 65536 > $O/r02_probe_lines.json 2> $O/probe.log
python scripts/bench_configs.py > $O/r02_configs.json 2> $O/configs.log
CPU_S=16 python scripts/bench_next_rows.py > $O/r02_next_rows.json 2> $O/next_rows.err
ls -la $O
