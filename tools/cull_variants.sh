# builds nothing: expects build_variants/lib_<CTAs per SM>_<unroll>.so = the library linked with rsc_cull.cu compiled with
# -DRSC_CULL_MINB=.. -DRSC_CULL_UNROLL=.. (RSC_LIB_PATH selects the library the Python host loads)
for v in 4_4 5_2 6_2 5_4 4_8 3_8; do echo -n "$v: "; RSC_LIB_PATH=$PWD/build_variants/lib_$v.so timeout 200 python tools/cull_bench.py --points 4194304 --reps 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['equal_counts'], min(d['culled_kernel_ms']))"; done
