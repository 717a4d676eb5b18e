L=spacecharge.jl_b200/lib
python tools/sort_probe.py f64 base
for v in s512k4 s256k8 s1024k8; do SCB_LIB=$L/libspacecharge_b200_$v.so python tools/sort_probe.py f64 $v; done
