o=gpurun_out; t=r2f
cap() { env RAST_LIB=$5 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -f -o $o/${t}_$3 python tools/quick_ab.py $4 --calls 2 > $o/${t}_ncu_$3.log 2>&1; }
cap k_resolve_shade 12 shade_w1 spin1080p ""
cap k_resolve_shade 12 shade_w4 spin1080p build/variants/librast_b200_w4.so
cap k_resolve_shade 12 shade_cache spin1080p build/variants/librast_b200_cache.so
cap k_raster_tiles 3 raster_tiles overdraw8k ""
cap k_setup 5 setup_pipe tess4k ""
