# ncu --set full of one call of the other BASELINE configs; only the one-line-per-launch summaries travel back
O=gpurun_out
for c in "$@"; do
  case $c in
    grid|hex) K='regex:bfs_smem_kernel'; N=2;;
    *) K='regex:bc_(forward|backward)(_look)?_kernel'; N=24;;
  esac
  ncu --set full --clock-control none -k "$K" -c $N -f -o /tmp/r02_cfg_$c python scripts/ncu_cfg.py $c > $O/cfg_$c.ncu.log 2>&1
  python scripts/ncu_brief.py /tmp/r02_cfg_$c.ncu-rep > $O/r02_ncu_cfg_$c.txt 2>/dev/null
  tail -n 1 $O/cfg_$c.ncu.log; wc -l $O/r02_ncu_cfg_$c.txt
done
