EMB_RC_DEBUG=1 EMB_RC_BATCH=1 timeout 300 python tools/rc_ab.py > gpurun_out/r2w_dbg.txt 2>&1
grep "max |Q" gpurun_out/r2w_dbg.txt | sort -t'|' -k3 -g | tail -3
grep "W u - Q R" gpurun_out/r2w_dbg.txt | awk '{print $(NF-4)}' | sort -g | tail -2
EMB_RC_BATCH=0 timeout 300 python tools/rc_ab.py 2>&1 | grep -v Warn
EMB_RC_BATCH=1 timeout 300 python tools/rc_ab.py 2>&1 | grep -v Warn
