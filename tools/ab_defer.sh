for nd in 0 1 0 1; do echo "nodefer $nd"; TTS_MEGA_NODEFER=$nd B_ONLY=1 timeout 120 python tools/quick_ar16.py 2>&1 | tail -1; done
TTS_MEGA_NODEFER=0 timeout 120 python tools/quick_ar16.py 2>&1 | tail -3
