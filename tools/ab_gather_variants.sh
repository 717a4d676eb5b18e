L=spacecharge.jl_b200/lib
python tools/variant_probe.py f64 base
SCB_CELL_GATHER=1 python tools/variant_probe.py f64 direct
for v in t128x3 m2 now m2x3; do SCB_LIB=$L/libspacecharge_b200_$v.so python tools/variant_probe.py f64 $v; done
SCB_RUNS_PER_SM=16 python tools/variant_probe.py f64 persm16
SCB_RUNS_PER_SM=256 python tools/variant_probe.py f64 persm256
python tools/variant_probe.py f32 base32
