cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q -k "topology or kuhn20 or c3d8 or c3d6 or shells or edge_cases or mixed_family or full_size_properties_c4 or region_growing or wide_node or library_loaded" 2>&1 | tail -8 > gpurun_out/r02_pytest11.log
python tools/topo_rate.py --n 220 >> gpurun_out/r02_topo11.log 2>&1
FEMB_TOPO_RADIX=1 python tools/topo_rate.py --n 220 >> gpurun_out/r02_topo11.log 2>&1
python tools/topo_rate.py --n 26 >> gpurun_out/r02_topo11.log 2>&1
FEMB_TOPO_RADIX=1 python tools/topo_rate.py --n 26 >> gpurun_out/r02_topo11.log 2>&1
cd $GRAFT_REPO_ROOT
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'bucket_|write_s|max_node|Device' -c 40 --csv --log-file gpurun_out/r02_topo_launches.csv python tools/topo_rate.py --n 220 > gpurun_out/r02_ncu12.log 2>&1
