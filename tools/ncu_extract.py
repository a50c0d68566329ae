"""Summarise an `ncu --set full` report: key metrics per captured launch -> JSON (profiles/).  usage: ncu_extract.py rep out.json"""
import csv, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = {'gpu__time_duration.sum': 'duration', 'launch__grid_size': 'grid', 'launch__block_size': 'block',
        'launch__registers_per_thread': 'regs_per_thread', 'launch__shared_mem_per_block_dynamic': 'dyn_smem',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active': 'tensor_pipe_active_pct_of_sm_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed': 'sm_throughput_pct',
        'dram__bytes_read.sum': 'dram_read', 'dram__bytes_write.sum': 'dram_write',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed': 'dram_throughput_pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed': 'l2_throughput_pct',
        'smsp__cycles_active.avg': 'smsp_cycles_active_avg', 'sm__cycles_elapsed.max': 'sm_cycles_elapsed_max',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum': 'lsu_global_store_sectors',
        'sm__warps_active.avg.pct_of_peak_sustained_active': 'achieved_occupancy_pct'}
launches = []
for r in rows[2:]:
    d = {'kernel': r[idx['Kernel Name']].split('(')[0]}
    for k, name in want.items():
        if k in idx:
            d[name] = f'{r[idx[k]]} {units[idx[k]]}'.strip()
    launches.append(d)
json.dump({'report': rep.split('/')[-1], 'launches': launches}, open(out, 'w'), indent=1)
print(f'{len(launches)} launches -> {out}')
