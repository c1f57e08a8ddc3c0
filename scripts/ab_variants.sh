# usage: bash scripts/ab_variants.sh [variant suffixes...]   ("" = librxcuda.so)
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_ab.log 2>&1
for v in "$@"; do
  [ "$v" = "default" ] && v=""
  L=/root/repo/rustworkx_b200/librxcuda$v.so
  RXCUDA_LIB=$L python scripts/ab_multi.py "S=16384" "S=1024" > gpurun_out/ab_ba$v.log 2>&1
  RXCUDA_LIB=$L GRAPH=gnm python scripts/ab_multi.py "S=4096;mode=edge" "S=4096" > gpurun_out/ab_gnm$v.log 2>&1
done
tail -n 3 gpurun_out/pytest_ab.log gpurun_out/ab_ba*.log gpurun_out/ab_gnm*.log
