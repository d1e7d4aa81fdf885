#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for f0 in 2 3 5; do RS_WIDE_STAGES_FWD0=$f0 python tools/c4_probe.py 1024 1000 2>&1 | tail -1 | sed "s/^/FWD0=$f0 /"; done
for f in 2 3 5 7; do RS_WIDE_STAGES_FWD=$f python tools/c4_probe.py 1024 1000 2>&1 | tail -1 | sed "s/^/FWD=$f /"; done
for b in 2 3 5; do RS_WIDE_STAGES_BWD=$b python tools/c4_probe.py 1024 1000 2>&1 | tail -1 | sed "s/^/BWD=$b /"; done
python tools/c4_probe.py 1024 1000 2>&1 | tail -1 | sed "s/^/default(4,4,4) /"
