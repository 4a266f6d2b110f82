"""Turn the ncu captures brought back in gpurun_out/ into the tracked summaries under profiles/.
    python scripts/summarise_profiles.py r01
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'
out_dir = os.path.join(REPO, 'profiles')
os.makedirs(out_dir, exist_ok=True)


def ms_of(v, u):
    v = float(v.replace(',', ''))
    return v / 1e6 if u in ('ns', 'nsecond') else v / 1e3 if u in ('us', 'usecond') else v * 1e3 if u in ('s', 'second') else v


def launch_list(path, title, note):
    lines = [ln for ln in open(path) if not ln.startswith('==')]
    tot, n = collections.OrderedDict(), 0
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        short = re.sub(r'\(.*', '', re.sub(r'<.*', '', row['Kernel Name'].replace('<unnamed>::', ''))).strip()[:70]
        d = tot.setdefault(short, [0, 0.0])
        d[0] += 1
        d[1] += ms_of(row['Metric Value'], row['Metric Unit'])
        n += 1
    total = sum(v[1] for v in tot.values())
    own = sum(v[1] for k, v in tot.items() if 'dg::' in k)
    md = ['# %s' % title, '', note, '',
          '%d launches, %.2f ms summed; own kernels (`dg::`) %.1f %% of it. Per-launch times under ncu are cold-cache '
          'and serialised: compare SHARES, not absolutes.' % (n, total, 100 * own / total), '',
          '| kernel | launches | ms | share |', '|---|---:|---:|---:|']
    for k, (c, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:32]:
        md.append('| `%s` | %d | %.3f | %.1f%% |' % (k, c, ms, 100 * ms / total))
    return '\n'.join(md) + '\n'


WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__occupancy_limit_registers', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'smsp__cycles_active.avg', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active']


def full_capture(rep, title, note):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [w for w in WANT if w in idx]
    md = ['# %s' % title, '', note, '', '| kernel | ' + ' | '.join(c.split('.')[0].replace('__', ' ') for c in cols) + ' |',
          '|---|' + '---:|' * len(cols)]
    recs = []
    for r in rows[2:]:
        name = re.sub(r'\(.*', '', r[idx['Kernel Name']])[:60]
        md.append('| `%s` | ' % name + ' | '.join('%s %s' % (r[idx[c]], units[idx[c]]) for c in cols) + ' |')
        recs.append({'kernel': name, **{c: (r[idx[c]], units[idx[c]]) for c in cols}})
    return '\n'.join(md) + '\n', recs


def gb(v, u):
    f = float(v.replace(',', ''))
    return f * {'byte': 1e-9, 'Kbyte': 1e-6, 'Mbyte': 1e-3, 'Gbyte': 1.0, 'Tbyte': 1e3}[u]


def raw_capture(path, title, note, keep=None):
    """Same table as full_capture, from a `--page raw --csv` export made on the GPU box (the .ncu-rep stays there)."""
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [w for w in WANT if w in idx]
    md = ['# %s' % title, '', note, '', '| # | kernel | ' + ' | '.join(c.split('.')[0].replace('__', ' ') for c in cols) + ' |',
          '|---:|---|' + '---:|' * len(cols)]
    recs = []
    for n, r in enumerate(rows[2:], 1):
        name = re.sub(r'\(.*', '', r[idx['Kernel Name']])
        name = re.sub(r'^void ', '', name).replace('dg::', '').replace('<unnamed>::', '')[:60]
        rec = {'n': n, 'kernel': name, **{c: (r[idx[c]], units[idx[c]]) for c in cols}}
        recs.append(rec)
        if keep is None or keep(rec):
            md.append('| %d | `%s` | ' % (n, name) + ' | '.join('%s %s' % (r[idx[c]], units[idx[c]]) for c in cols) + ' |')
    return '\n'.join(md) + '\n', recs


g = os.path.join(REPO, 'gpurun_out')
import shutil
final = os.path.isfile(os.path.join(g, 'launches_final.csv'))
if final:
    bench = json.loads(open(os.path.join(g, 'bench_final.json')).read().strip().splitlines()[-1])
    open(os.path.join(out_dir, '%s_launches_syn20m_step.md' % tag), 'w').write(launch_list(
        os.path.join(g, 'launches_final.csv'), 'Round 1 -- ncu launch list of ONE training step at syn20m (final state of the round)',
        'Command (gpurun, 1x B200): `DG_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum '
        '--clock-control none --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline '
        '--eager` (bench.py had exited 0 without ncu directly before; `--eager` = per-kernel launches, the default run replays '
        'the same kernels from one CUDA graph). The same step timed with CUDA events without ncu: %.1f ms (graph replay), '
        '~%.1f ms eager.' % (bench['ms_per_step'], bench['ms_per_step'] + 1.0)))
    shutil.copy(os.path.join(g, 'launches_final.csv'), os.path.join(out_dir, '%s_launches_syn20m_step.csv' % tag))
    shutil.copy(os.path.join(g, 'bench_final.json'), os.path.join(out_dir, '%s_bench_syn20m.json' % tag))
    md, recs = raw_capture(os.path.join(g, 'spmm_final_raw.csv'),
                           'Round 1 -- ncu --set full of all 30 spmm_csr_kernel launches of one syn20m training step',
                           'Command: `DG_PROFILE_RANGE=1 ncu --profile-from-start off --set full --clock-control none -k regex:spmm_csr '
                           '-c 30 -o /tmp/spmm_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eager`, exported on the box '
                           'with `ncu -i ... --page raw --csv`. Launch order: 1-2 GCMC layer 0 forward (d=344; dst=drug, dst=disease), 3-6 '
                           'GCMC layers 1-2 forward (d=128), 7-14 FGCN forward (d=768 / 128 over the four kNN graphs), 15-16 decoder '
                           'segment sums (d=128, one row per scored pair: pure HBM streaming), 17-24 FGCN backward, 25-28 GCMC layers 2-1 '
                           'backward, 29-30 GCMC layer 0 backward (transposed relation blocks).')
    open(os.path.join(out_dir, '%s_spmm_ncu_full.md' % tag), 'w').write(md)
    tot = lambda r: gb(*r['dram__bytes_read.sum']) + gb(*r['dram__bytes_write.sum'])
    by = {'gcmc_d344': [1, 2, 29, 30], 'gcmc_d128': [3, 4, 5, 6, 25, 26, 27, 28], 'fgcn_d768': [7, 11, 20, 24],
          'decoder_d128': [15, 16]}
    traffic = {k: round(sum(tot(recs[i - 1]) for i in v) / len(v) * 1e9) for k, v in by.items()}
    traffic['_note'] = ('dram__bytes_read.sum + dram__bytes_write.sum per launch (bytes), mean over the launches of the class in one '
                        'training step (forward and backward); from profiles/%s_spmm_ncu_full.md' % tag)
    json.dump(traffic, open(os.path.join(out_dir, 'roofline_traffic.json'), 'w'), indent=1)
    md, _ = raw_capture(os.path.join(g, 'tc_final_raw.csv'),
                        'Round 1 -- ncu --set full of the tcgen05 kernels (projection GEMM, decoder forward / backward) in a syn20m step',
                        'Command: `... ncu --set full --clock-control none -k regex:"gemm_nt_tf32|decoder_fwd_tc|decoder_bwd_tc" -c 12 ...` on the '
                        'same bench command. `sm pipe_tensor_cycles_active` is the tensor-pipe utilisation; the GEMM rows are the first '
                        'projections of the step (GCMC layer 0: [100k x 1024] . [1024 x 344] x 2 relations, then the disease side).')
    if os.path.isfile(os.path.join(g, 'dec_final_raw.csv')):
        md2, _ = raw_capture(os.path.join(g, 'dec_final_raw.csv'), 'tcgen05 decoder kernels at the syn20m pair count (20 M pairs)',
                             'Command: `ncu --set full --clock-control none -k regex:decoder_.*tc -c 4 python scripts/decoder_check.py '
                             '--only-time --time` (label order, p = 0; rows alternate forward / backward).')
        md += '\n' + md2
    md += ('\nSASS of the two tensor-core objects (`cuobjdump -sass dreamgnn_b200/build/{gemm_tc,decoder_tc}.o`): gemm_tc 48 x UTCHMMA, '
           '48 x UTMALDG.3D, 4 x LDTM.x32, 8 x UTCBAR; decoder_tc 144 x UTCHMMA, 13 x LDTM.x16, 4 x UTCBAR (operands written by the '
           'gather threads, no TMA).\n')
    open(os.path.join(out_dir, '%s_gemm_decoder_ncu_full.md' % tag), 'w').write(md)
    print(traffic)
    sys.exit(0)
g = os.path.join(REPO, 'gpurun_out')
if os.path.isfile(os.path.join(g, 'launches3.csv')):
    open(os.path.join(out_dir, '%s_launches_syn20m_step.md' % tag), 'w').write(launch_list(
        os.path.join(g, 'launches3.csv'), 'Round 1 -- ncu launch list of ONE training step at syn20m (final state of the round)',
        'Command (gpurun, 1x B200): `DG_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum '
        '--clock-control none --csv --log-file gpurun_out/launches3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline` '
        '(the same command had exited 0 without ncu directly before). The same step timed with CUDA events without ncu: 102.5 ms.'))
    import shutil
    shutil.copy(os.path.join(g, 'launches3.csv'), os.path.join(out_dir, '%s_launches_syn20m_step.csv' % tag))
traffic = {}
if os.path.isfile(os.path.join(g, 'spmm_full2.ncu-rep')):
    md, recs = full_capture(os.path.join(g, 'spmm_full2.ncu-rep'), 'Round 1 -- ncu --set full of the first six spmm_csr_kernel launches of a syn20m step',
                            'Command: `DG_PROFILE_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on '
                            '-k regex:spmm_csr -c 6 -o gpurun_out/spmm_full2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline`. '
                            'Launches 1-2: GCMC layer 0 forward (d=344; dst=drug then dst=disease); 3-6: layers 1-2 forward (d=128).')
    open(os.path.join(out_dir, '%s_spmm_ncu_full.md' % tag), 'w').write(md)
    by_d = collections.defaultdict(list)
    for r in recs:
        d = 344 if 'LoadF32, 3' in r['kernel'] or ', 3, ' in r['kernel'] else 128
        by_d[d].append(gb(*r['dram__bytes_read.sum']) + gb(*r['dram__bytes_write.sum']))
    for d, v in by_d.items():
        traffic['gcmc_d%d' % d] = round(sum(v) / len(v) * 1e9)
if os.path.isfile(os.path.join(g, 'gemm_dec_full.ncu-rep')):
    md, recs = full_capture(os.path.join(g, 'gemm_dec_full.ncu-rep'), 'Round 1 -- ncu --set full of the tcgen05 GEMM and decoder kernels (syn20m step)',
                            'Command: `... ncu --set full -k regex:"gemm_nt_tf32|decoder_fwd|decoder_bwd" -c 4 ...` on the same bench command.')
    open(os.path.join(out_dir, '%s_gemm_decoder_ncu_full.md' % tag), 'w').write(md)
if traffic:
    traffic['_note'] = 'dram__bytes_read.sum + dram__bytes_write.sum per launch (bytes), mean over the captured forward launches; from profiles/%s_spmm_ncu_full.md' % tag
    json.dump(traffic, open(os.path.join(out_dir, 'roofline_traffic.json'), 'w'), indent=1)
if os.path.isfile(os.path.join(g, 'bench_r01.json')):
    import shutil
    shutil.copy(os.path.join(g, 'bench_r01.json'), os.path.join(out_dir, '%s_bench_syn20m.json' % tag))
print(os.listdir(out_dir))
