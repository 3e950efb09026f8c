for v in 0 1 0 1; do echo "noevict $v"; TTS_MEGA_NOEVICT=$v B_ONLY=1 timeout 120 python tools/quick_ar16.py 2>&1 | tail -1; done
