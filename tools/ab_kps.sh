for k in 32 64 128 32 128; do echo "kps $k"; PER_STEP=1 TTS_MEGA_KPS=$k B_ONLY=1 timeout 120 python tools/quick_ar16.py 2>&1 | tail -2; done
