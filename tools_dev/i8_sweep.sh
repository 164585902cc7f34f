args=""
for R in 2 3 4 6; do for k in 2 3 4 6 8 12; do args="$args DCG_I8_RING=$R,DCG_I8_WINDOW_STAGES=$k,DCG_I8_RING_BYTES=200000000"; done; done
timeout 400 python tools_dev/i8_fused_time.py $args 2>&1 | tail -30
