# GPU utilisation / memory logger (reference: /root/reference/statistics.sh) - one CSV per entrypoint, 500 ms period.
# usage: sh statistics.sh <name>   (writes <name>_log.csv until killed)
nvidia-smi --query-gpu=timestamp,index,memory.total,memory.used,memory.free,utilization.gpu,utilization.memory,clocks.sm,power.draw --format=csv -lms 500 -f ${1:-distributed}_log.csv
