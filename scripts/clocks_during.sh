#!/bin/bash
# sample SM clocks / power / throttle reasons while a command runs: clocks_during.sh out.csv cmd...
out=$1; shift
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv -lms 100 > $out &
pid=$!
"$@"
kill $pid
