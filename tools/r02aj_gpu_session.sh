Q=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap
for ms in 0 100 200 500; do
  echo "sampler -lms $ms" >> gpurun_out/r02aj_sampler_effect.log
  if [ $ms -gt 0 ]; then nvidia-smi --query-gpu=$Q --format=csv,noheader,nounits -lms $ms -i 0 > /dev/null 2>&1 & SP=$!; fi
  python tools/step_times.py n2d2 dvbs2 10 post 2>&1 | grep "profile mode 0" >> gpurun_out/r02aj_sampler_effect.log
  if [ $ms -gt 0 ]; then kill $SP; fi
done
cat gpurun_out/r02aj_sampler_effect.log
