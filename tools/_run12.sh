cd $GRAFT_REPO_ROOT
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'bucket_|write_s|max_node|Device' -c 40 --csv --log-file gpurun_out/r02_topo_launches.csv python tools/topo_rate.py --n 220 > gpurun_out/r02_ncu12.log 2>&1
