set -x
EMB_RC_BATCH=0 timeout 300 python tools/rc_ab.py 2>&1 | grep -v Warn
EMB_RC_BATCH=1 timeout 300 python tools/rc_ab.py 2>&1 | grep -v Warn
