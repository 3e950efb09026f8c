for rep in 8 4 2 1; do for v2 in 0 1; do echo "rep $rep v2 $v2"; TTS_MEGA_REP=$rep TTS_MEGA_V2=$v2 B_ONLY=1 timeout 120 python tools/quick_ar16.py 2>&1 | tail -1; done; done
