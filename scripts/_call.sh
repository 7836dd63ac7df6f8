mkdir -p gpurun_out
out=gpurun_out/r2b_tile_r3_c4.jsonl
: > $out
timeout 600 python scripts/jit_tile_sweep.py c4 "9,3,96,4" 2>> gpurun_out/r2b_r3.err >> $out
QSX_JIT_MINCTAS=2 timeout 600 python scripts/jit_tile_sweep.py c4 "9,3,96,3" 2>> gpurun_out/r2b_r3.err >> $out
QSX_JIT_MINCTAS=4 timeout 600 python scripts/jit_tile_sweep.py c4 "8,3,96,3" 2>> gpurun_out/r2b_r3.err >> $out
QSX_JIT_MINCTAS=3 timeout 600 python scripts/jit_tile_sweep.py c4 "8,3,96,3" 2>> gpurun_out/r2b_r3.err >> $out
cut -c1-330 $out
