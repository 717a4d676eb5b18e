for z in 0 2 4 8 16 32; do SCB_ZCHUNK=$z python bench.py --stages-only 2>/dev/null | sed "s/^/zchunk=$z /"; done
for z in 0 4 8; do SCB_ZCHUNK=$z python bench.py --stages-only --dtype f32 2>/dev/null | sed "s/^/zchunk=$z /"; done
for z in 0 4 8; do SCB_ZCHUNK=$z python bench.py --stages-only --workload cathode 2>/dev/null | sed "s/^/zchunk=$z /"; done
