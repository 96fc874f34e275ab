# same-box A/B of differently built libraries: PYGPR_B200_LIB selects the .so (pygpr_b200/_lib.py)
for lib in "$@"; do
echo "== $lib"
PYGPR_B200_LIB=$PWD/$lib python tools/perf_probe.py 148,8192,16,1024 2>&1 | tail -2
PYGPR_B200_LIB=$PWD/$lib python tools/grbcm_share.py 32 2>&1 | grep -E "overlapped|lml"
done
