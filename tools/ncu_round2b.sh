set -x
ncu --graph-profiling graph --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02b_graph_step.csv python tools/ncu_stream.py graph > gpurun_out/r02b_graph.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches_standalone.csv python tools/ncu_stream.py standalone > gpurun_out/r02b_standalone.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_recurrence_stream -c 8 -o gpurun_out/r02b_prof_rec -f python tools/ncu_stream.py standalone > gpurun_out/r02b_full.log 2>&1
tail -2 gpurun_out/r02b_graph.log gpurun_out/r02b_standalone.log gpurun_out/r02b_full.log
ls -la gpurun_out/
