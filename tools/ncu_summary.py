"""Summarise an .ncu-rep (raw page) into the handful of numbers we track per kernel."""
import csv, subprocess, sys, io, json
KEYS = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum',
 'l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active',
 'launch__registers_per_thread','launch__grid_size','launch__block_size','smsp__inst_executed.sum',
 'smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__data_pipe_lsu_wavefronts.sum',
 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum',
 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
 'sm__cycles_elapsed.max','smsp__cycles_active.avg','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_fma.sum','sm__inst_executed_pipe_alu.sum',
 'sm__inst_executed_pipe_lsu.sum','sm__inst_executed_pipe_xu.sum','smsp__inst_executed_op_global_red.sum',
 'l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum','lts__t_sectors_op_red.sum','lts__t_sectors_op_atom.sum']
def main(path):
    out = subprocess.run(['ncu','-i',path,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {'kernel': r[hdr.index('Kernel Name')][:90]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k); d[k] = r[i] + ' ' + units[i]
        st = {hdr[i].replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''): float(r[i] or 0)
              for i in range(len(hdr)) if hdr[i].startswith('smsp__average_warps_issue_stalled_') and hdr[i].endswith('_per_issue_active.ratio')}
        d['stalls_per_issue'] = dict(sorted(st.items(), key=lambda kv: -kv[1])[:8])
        res.append(d)
    print(json.dumps(res, indent=1))
if __name__ == '__main__':
    main(sys.argv[1])
