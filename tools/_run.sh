python tools/prof_packed.py > gpurun_out/plain_packed.log 2>&1 && tail -1 gpurun_out/plain_packed.log &&
ncu --clock-control none -k k_tridiag_blk -s 5 -c 1 --metrics gpu__time_duration.sum,l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum,l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --csv --log-file gpurun_out/r2_tma_packed.csv python tools/prof_packed.py > gpurun_out/ncu_packed.log 2>&1
tail -8 gpurun_out/r2_tma_packed.csv | cut -c1-260
