mkdir -p gpurun_out
nvidia-smi --id=0 --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv,noheader,nounits -lms 100 > gpurun_out/smi_test.txt 2> gpurun_out/smi_test.err &
PID=$!
sleep 1.5
kill $PID
echo "--- rows:"; head -5 gpurun_out/smi_test.txt; echo "--- err:"; head -5 gpurun_out/smi_test.err
