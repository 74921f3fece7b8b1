"""Turn the gpurun_out/ ncu captures and bench lines into the committed summaries under profiles/."""
import collections, csv, json, shutil
keep=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__waves_per_multiprocessor','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','gpc__cycles_elapsed.max','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','smsp__pcsamp_warps_issue_stalled_wait','smsp__pcsamp_warps_issue_stalled_math_pipe_throttle','smsp__pcsamp_warps_issue_stalled_barrier','smsp__pcsamp_warps_issue_stalled_short_scoreboard','smsp__pcsamp_warps_issue_stalled_long_scoreboard','smsp__pcsamp_warps_issue_stalled_selected','smsp__pcsamp_warps_issue_stalled_not_selected','smsp__pcsamp_warps_issue_stalled_no_instructions','smsp__pcsamp_warps_issue_stalled_mio_throttle']
CMD='python bench.py --steps 2 --warmup 3 --e2e-steps 1 --cpu-budget 0 --t0-sample 2368 --no-check'
for k,name in (('t1','t1_fused_kernel'),('t0','t0_fused_wpm_kernel'),('base','base_inverse_kernel')):
    rows=list(csv.reader(open(f'gpurun_out/r01_{k}_raw.csv')))
    hdr,units,vals=rows[0],rows[1],rows[2]
    with open(f'profiles/r01_{name}_ncu_full_excerpt.csv','w') as f:
        f.write(f'# ncu --set full --clock-control none, one launch of {name} inside: python scratch/one_call.py c4 ' + ('lu' if k == 't0' else 'auto') + ' 1\n')
        kn = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''
        f.write(f'# kernel: {kn}\nmetric,unit,value\n')
        for h,u,v in zip(hdr,units,vals):
            if h in keep: f.write(f'{h},{u},{v}\n')
shutil.copy('gpurun_out/r01_launches.csv','profiles/r01_launches_bench_c4.csv')
for f in ['r01_bench_c4','r01_bench_c4_reference','r01_bench_c5_lu','r01_bench_c5_lu_k32','r01_bench_c5_lu_k100','r01_bench_c5_lu_k316','r01_bench_c3','r01_bench_c2']:
    json.load(open(f'gpurun_out/{f}.json'))
    shutil.copy(f'gpurun_out/{f}.json', f'profiles/{f}.json')
rows = list(csv.reader(open('gpurun_out/r01_launches.csv')))
hdr_i = next(i for i,r in enumerate(rows) if r and r[0]=='ID')
hdr = rows[hdr_i]
ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[hdr_i+1:]:
    if len(r) <= vi: continue
    name = r[ki].split('(')[0][:60]; val = float(r[vi].replace(',','')); unit = r[ui]
    if unit == 'ns': val/=1e3
    elif unit == 'ms': val*=1e3
    a = agg.setdefault(name, [0,0.0]); a[0]+=1; a[1]+=val
with open('profiles/r01_launches_summary.txt','w') as f:
    f.write(f"ncu --metrics gpu__time_duration.sum --clock-control none, command: {CMD}\n")
    f.write("(5 device-resident steps + 2 e2e calls (each with a device Newton-Schulz inverse: the many gemm/ns_residual launches) + T0 sample leg + closed-form leg + FP64 peak probes; cold-cache serialised times: compare shares)\n\n")
    f.write(f"{'launches':>8} {'total us':>12} {'avg us':>10}  kernel\n")
    for k,v in sorted(agg.items(), key=lambda x:-x[1][1]):
        f.write(f"{v[0]:8d} {v[1]:12.1f} {v[1]/v[0]:10.1f}  {k}\n")
print(open('profiles/r01_launches_summary.txt').read()[:2200])
# DRAM traffic of the dominant kernel per launch, read by bench.py for the roofline line
tr = {}
rows=list(csv.reader(open('gpurun_out/r01_t1_raw.csv'))); d=dict(zip(rows[0],rows[2])); u=dict(zip(rows[0],rows[1]))
def to_bytes(key):
    v=float(d[key].replace(',','')); unit=u[key].lower()
    return v*{'byte':1,'kbyte':1e3,'mbyte':1e6,'gbyte':1e9}[unit]
tr['c4:n1:tier1']={'kernel':'t1_fused_kernel','dram_bytes_read':to_bytes('dram__bytes_read.sum'),'dram_bytes_write':to_bytes('dram__bytes_write.sum'),'source':'profiles/r01_t1_fused_kernel_ncu_full_excerpt.csv'}
json.dump(tr, open('profiles/traffic.json','w'), indent=1)
print(tr)
for k in ('t1_fused_kernel','t0_fused_wpm_kernel','base_inverse_kernel'):
    print(open(f'profiles/r01_{k}_ncu_full_excerpt.csv').read())
for f in ['r01_bench_c4','r01_bench_c5_lu','r01_bench_c3','r01_bench_c2']:
    d=json.load(open(f'profiles/{f}.json'))
    print(f, "%.3e"%d['value'], round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_call'],3), 'frac', round(d['roofline']['frac'],4), d['parity']['max_rel_err'], {k:(round(v.get('achieved_tflops',0),2), round(v.get('frac_of_fp64_peak',0),4)) if 'achieved_tflops' in v else v.get('ms_per_matrix_stages_2_3') for k,v in d['kernels'].items()})
