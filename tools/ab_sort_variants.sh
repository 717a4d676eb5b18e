L=spacecharge.jl_b200/lib
python tools/sort_probe.py f64 base
for v in k16; do SCB_LIB=$L/libspacecharge_b200_$v.so python tools/sort_probe.py f64 $v; done
python tools/sort_probe.py f32 base32
