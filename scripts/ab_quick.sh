# usage: bash scripts/ab_quick.sh suffix...   -- BA S=8192 only, no tests
for v in "$@"; do
  [ "$v" = "default" ] && v=""
  L=/root/repo/rustworkx_b200/librxcuda$v.so
  echo "== $v" >> gpurun_out/ab_quick.log
  RXCUDA_LIB=$L REPS=1 python scripts/ab_multi.py "S=8192" >> gpurun_out/ab_quick.log 2>&1
done
cat gpurun_out/ab_quick.log
