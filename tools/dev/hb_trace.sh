#!/bin/bash
# developer: in-graph interval from the head-backward kernel's start to the next main-stream kernel's start, per experiment
for ex in 0 1 2; do
  SB_TC_HEAD_BWD=1 SB_HB_EXPERIMENT=$ex python tools/trace_step.py cfg2 64 2>/dev/null | grep -E "^# cfg2|\(242, 192\)|\(1, 352\)" | tr '\n' ' '; echo " [exp $ex]"
done
SB_TC_HEAD_BWD=0 python tools/trace_step.py cfg2 64 2>/dev/null | grep -E "^# cfg2|\(484, 128\)|\(1, 352\)" | tr '\n' ' '; echo " [ffma]"
