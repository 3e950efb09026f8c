for sp in 0 150 400 800 0 150; do echo "spin $sp"; TTS_MEGA_SPIN=$sp B_ONLY=1 timeout 120 python tools/quick_ar16.py 2>&1 | tail -1; done
