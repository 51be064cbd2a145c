# [historical: produced profiles/r2_ab_narrow_phase.txt; the DOPER_OLD_NARROW switch and the sqrt-free narrow phase it compared
#  against were removed again, see DESIGN.md section 5.2c]
for wl in "c5 0" "c4 0" "c2 4096" "c2 32768"; do
  for lib in doper_b200/_lib/libdoper_b200_oldnarrow.so doper_b200/_lib/libdoper_b200.so; do
    echo "== $wl $lib"; DOPER_B200_LIB=$PWD/$lib timeout 300 python tools/tune_baked.py $wl 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:v[0] for k,v in d.items() if isinstance(v,list)})"
  done
done
