for lib in pygpr_b200/libpygpr_b200.so tools/libpygpr_b200_s5.so; do
echo "== $lib"
PYGPR_B200_LIB=$PWD/$lib python tools/perf_probe.py 148,8192,16,1024 2>&1 | tail -2
PYGPR_B200_LIB=$PWD/$lib python tools/grbcm_share.py 32 2>&1 | grep -E "overlapped|lml"
done
