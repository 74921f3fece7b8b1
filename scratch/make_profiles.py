"""Turn the gpurun_out/ ncu captures and bench lines into the committed summaries under profiles/ (round 2)."""
import collections, csv, json, shutil
R = 'r02'
keep=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__waves_per_multiprocessor','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__shared_mem_per_block_dynamic','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','lts__t_bytes.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','gpc__cycles_elapsed.max','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','smsp__pcsamp_warps_issue_stalled_wait','smsp__pcsamp_warps_issue_stalled_math_pipe_throttle','smsp__pcsamp_warps_issue_stalled_barrier','smsp__pcsamp_warps_issue_stalled_short_scoreboard','smsp__pcsamp_warps_issue_stalled_long_scoreboard','smsp__pcsamp_warps_issue_stalled_selected','smsp__pcsamp_warps_issue_stalled_not_selected','smsp__pcsamp_warps_issue_stalled_no_instructions','smsp__pcsamp_warps_issue_stalled_mio_throttle','smsp__pcsamp_warps_issue_stalled_branch_resolving']
CMD='python bench.py --steps 2 --warmup 3 --e2e-steps 1 --cpu-budget 0 --t0-sample 2368 --no-check'
caps = (('t1','t1_fused_kernel','c4 auto 1'),('t0n100','t0_fused_ll_kernel_n100','c4 lu 1'),('t0n64','t0_fused_ll_kernel_n64','c5 lu 1'))
tr = {}
def to_bytes(d, u, key):
    v=float(d[key].replace(',','')); unit=u[key].lower()
    return v*{'byte':1,'kbyte':1e3,'mbyte':1e6,'gbyte':1e9}[unit]
for k,name,args in caps:
    rows=list(csv.reader(open(f'gpurun_out/{R}_{k}_raw.csv')))
    hdr,units,vals=rows[0],rows[1],rows[2]
    d=dict(zip(hdr,vals)); u=dict(zip(hdr,units))
    with open(f'profiles/{R}_{name}_ncu_full_excerpt.csv','w') as f:
        f.write(f'# ncu --set full --clock-control none, one launch inside: python scratch/one_call.py {args}' + (' (4096 pairs: 64 bra x 64 ket excitations)' if 'lu' in args else '') + '\n')
        kn = d.get('Kernel Name','')
        f.write(f'# kernel: {kn}\nmetric,unit,value\n')
        for h,uu,v in zip(hdr,units,vals):
            if h in keep: f.write(f'{h},{uu},{v}\n')
    tr[{'t1':'c4:n1:tier1','t0n100':'c4:n1:tier0:4096pairs','t0n64':'c5:n1:tier0:4096pairs'}[k]] = {
        'kernel': kn.split('(')[0], 'dram_bytes_read': to_bytes(d,u,'dram__bytes_read.sum'), 'dram_bytes_write': to_bytes(d,u,'dram__bytes_write.sum'),
        'source': f'profiles/{R}_{name}_ncu_full_excerpt.csv'}
json.dump(tr, open('profiles/traffic.json','w'), indent=1)
print(json.dumps(tr, indent=1))
shutil.copy(f'gpurun_out/{R}_launches.csv',f'profiles/{R}_launches_bench_c4.csv')
benches = [f'{R}_bench_c4',f'{R}_bench_c4_reference',f'{R}_bench_c5_lu',f'{R}_bench_c5_lu_k32',f'{R}_bench_c5_lu_k100',f'{R}_bench_c5_lu_k316',f'{R}_bench_c3',f'{R}_bench_c3_lu',f'{R}_bench_c2',f'{R}_bench_traj',
           f'{R}_scale_1',f'{R}_scale_2',f'{R}_scale_4',f'{R}_scale_8',f'{R}_scale_8_nccl',f'{R}_scale_c5_lu_1',f'{R}_scale_c5_lu_8']
for f in benches:
    line = open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1]
    json.loads(line)
    open(f'profiles/{f}.json','w').write(line + '\n')
rows = list(csv.reader(open(f'gpurun_out/{R}_launches.csv')))
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
with open(f'profiles/{R}_launches_summary.txt','w') as f:
    f.write(f"ncu --metrics gpu__time_duration.sum --clock-control none, command: {CMD}\n")
    f.write("(5 device-resident steps + 2 e2e calls (each with a device Newton-Schulz inverse: the gemm/ns_residual launches) + T0 sample leg + closed-form leg + FP64 peak probes; cold-cache serialised times: compare shares)\n\n")
    f.write(f"{'launches':>8} {'total us':>12} {'avg us':>10}  kernel\n")
    for k,v in sorted(agg.items(), key=lambda x:-x[1][1]):
        f.write(f"{v[0]:8d} {v[1]:12.1f} {v[1]/v[0]:10.1f}  {k}\n")
print(open(f'profiles/{R}_launches_summary.txt').read()[:2500])
for k,name,args in caps:
    print(open(f'profiles/{R}_{name}_ncu_full_excerpt.csv').read())
