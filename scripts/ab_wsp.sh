for v in "$@"; do
  [ "$v" = "default" ] && v=""
  echo "== $v"
  RXCUDA_LIB=/root/repo/rustworkx_b200/librxcuda$v.so CPU_S=16 python scripts/bench_next_rows.py 2>/dev/null | grep newman | cut -c1-200
done
