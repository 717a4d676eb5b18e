L=spacecharge.jl_b200/lib
python tools/variant_probe.py f64 base
for v in pf2 nopf pf256; do SCB_LIB=$L/libspacecharge_b200_$v.so python tools/variant_probe.py f64 $v; done
python tools/variant_probe.py f32 base32
SCB_LIB=$L/libspacecharge_b200_pf2.so python tools/variant_probe.py f32 pf2_32
