for v in "$@"; do
  [ "$v" = "default" ] && v=""
  echo "== $v" >> gpurun_out/ab_gnm_quick.log
  RXCUDA_LIB=/root/repo/rustworkx_b200/librxcuda$v.so GRAPH=gnm python scripts/ab_multi.py "S=4096;mode=edge" >> gpurun_out/ab_gnm_quick.log 2>&1
done
cat gpurun_out/ab_gnm_quick.log
