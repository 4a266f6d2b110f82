"""Round-2 profiles: turns the ncu CSV exports brought back in gpurun_out/ (scripts/r02_capture.sh) into tracked summaries
under profiles/, and stamps profiles/roofline_traffic.json with the digest of the SpMM sources the capture was taken from
(bench.py ignores a capture whose digest does not match the loaded build).
    python scripts/summarise_r02.py
"""
import collections
import csv
import json
import os
import re
import shutil

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(REPO, 'gpurun_out'), os.path.join(REPO, 'profiles')
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}


def short(name):
    name = re.sub(r'\(.*', '', name)
    return re.sub(r'^void ', '', name).replace('dg::', '').replace('<unnamed>::', '').replace('unnamed>::', '')[:58]


def ms_of(v, u):
    v = float(v.replace(',', ''))
    return v / 1e6 if u.startswith('n') else v / 1e3 if u.startswith('u') else v * 1e3 if u.startswith('s') else v


# ---- (1) targeted metrics of every own kernel of one eager step ------------------------------------------------
path = os.path.join(G, 'r02_own_kernels_metrics.csv')
if os.path.isfile(path):
    rows = collections.OrderedDict()
    for x in csv.DictReader([ln for ln in open(path) if ln.startswith('"')]):
        d = rows.setdefault(int(x['ID']), {'name': short(x['Kernel Name']), 'grid': x['Grid Size']})
        d[x['Metric Name']] = (x['Metric Value'], x['Metric Unit'])
    md = ['# Round 2 -- memory metrics of every own kernel in ONE eager syn20m training step', '',
          'Command (gpurun, 1x B200): `DG_PROFILE_RANGE=1 ncu --profile-from-start off --metrics dram__bytes_read.sum,'
          'dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__throughput...,gpu__dram_throughput...,'
          'sm__pipe_tensor_cycles_active...,launch__registers_per_thread,launch__grid_size --clock-control none -k regex:"spmm_csr|'
          'decoder_|gemm_nt|colsum_partial|splitk_reduce|center_normalize|compact_write" --csv python bench.py --steps 1 --warmup 3 '
          '--no-cpu-baseline --eager` (scripts/r02_capture.sh). Times under ncu are per kernel in isolation (cold L2, no '
          'neighbours): the in-step durations of the SpMM launches (CUDA events inside the real step) are in '
          '`r02_bench_syn20m.json: spmm_launches`. SpMM instances ending in `, 1>` are the L2-prefetch ones.', '',
          '| # | kernel | grid | ms | dram read MB | dram write MB | L2 hit % | dram % | lts % | tensor % | regs |',
          '|---:|---|---|---:|---:|---:|---:|---:|---:|---:|---:|']
    recs = []
    for k, d in rows.items():
        g = lambda m: d.get(m, ('0', ''))
        b = lambda m: float(g(m)[0].replace(',', '')) * UNIT.get(g(m)[1], 1.0)
        rec = dict(id=k, name=d['name'], grid=d['grid'], ms=ms_of(*g('gpu__time_duration.sum')), rd=b('dram__bytes_read.sum'),
                   wr=b('dram__bytes_write.sum'))
        recs.append(rec)
        md.append('| %d | `%s` | %s | %.3f | %.1f | %.1f | %s | %s | %s | %s | %s |' % (
            k, d['name'], d['grid'], rec['ms'], rec['rd'] / 1e6, rec['wr'] / 1e6, g('lts__t_sector_hit_rate.pct')[0][:5],
            g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')[0][:5],
            g('lts__throughput.avg.pct_of_peak_sustained_elapsed')[0][:5],
            g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')[0][:5], g('launch__registers_per_thread')[0]))
    open(os.path.join(P, 'r02_own_kernels_metrics.md'), 'w').write('\n'.join(md) + '\n')
    # DRAM traffic per launch of each SpMM class (bench.py reads this for roofline.traffic)
    spmm = [r for r in recs if 'spmm_csr' in r['name']]
    cls = collections.defaultdict(list)
    for r in spmm:
        wide = 'LoadF32, 3' in r['name'] or 'stage_kernel<3' in r['name']     # d = 344 / 768: the staged instance since round 2
        seg = 'LoadF32, 1, 8, 0' in r['name']            # unweighted d=128: the decoder's segment sums
        if seg:                                              # unweighted d=128: the decoder's node-gradient sums
            cls['decoder_seg_d128' if r['ms'] > 0.8 else 'decoder_slots_d128'].append(r)
        elif wide and r['ms'] > 1.0:
            cls['gcmc_d344'].append(r)
        elif wide:
            cls['fgcn_d768'].append(r)
        elif r['ms'] > 0.5:
            cls['gcmc_d128'].append(r)
        else:
            cls['fgcn_d128'].append(r)
    traffic = {k: round(sum(r['rd'] + r['wr'] for r in v) / len(v)) for k, v in cls.items()}
    traffic['_launches'] = {k: len(v) for k, v in cls.items()}
    import sys
    sys.path.insert(0, REPO)
    from dreamgnn_b200 import build as _build
    traffic['kernel_digest'] = _build.kernel_digest(['spmm.cu', 'common.cuh'])
    traffic['source'] = 'profiles/r02_own_kernels_metrics.md (ncu dram__bytes_read.sum + dram__bytes_write.sum, one eager syn20m step)'
    traffic['_note'] = ('dram__bytes_read.sum + dram__bytes_write.sum per launch (bytes), mean over the launches of the class in one '
                        'training step (forward and backward); from profiles/r02_own_kernels_metrics.md')
    json.dump(traffic, open(os.path.join(P, 'roofline_traffic.json'), 'w'), indent=1)
    print(traffic)

# ---- (2) --set full raw exports --------------------------------------------------------------------------------
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__occupancy_limit_registers', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.pct', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio']
for fn, title in (('r02_spmm_d344_full_raw.csv', 'ncu --set full of the dominant SpMM class (GCMC layer 0 forward, d=344, L2-prefetch instance)'),
                  ('r02_decoder_full_raw.csv', 'ncu --set full of the tcgen05 decoder kernels (forward with 256-bit z2 stores, backward)')):
    path = os.path.join(G, fn)
    if not os.path.isfile(path):
        continue
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [w for w in WANT if w in idx]
    md = ['# Round 2 -- ' + title, '',
          'Command: `DG_PROFILE_RANGE=1 ncu --profile-from-start off --set full --clock-control none -k regex:<kernel> -c 2 -o /tmp/x '
          'python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eager`, exported on the box with `ncu -i /tmp/x.ncu-rep --page raw --csv` '
          '(scripts/r02_capture.sh).', '', '| metric | ' + ' | '.join('launch %d' % (i + 1) for i in range(len(rows) - 2)) + ' |',
          '|---|' + '---:|' * (len(rows) - 2)]
    md.append('| kernel | ' + ' | '.join('`%s`' % short(r[idx['Kernel Name']]) for r in rows[2:]) + ' |')
    for c in cols:
        md.append('| %s [%s] | ' % (c, units[idx[c]]) + ' | '.join(r[idx[c]] for r in rows[2:]) + ' |')
    open(os.path.join(P, fn.replace('_raw.csv', '.md')), 'w').write('\n'.join(md) + '\n')

# ---- (3) launch list + bench lines -------------------------------------------------------------------------------
path = os.path.join(G, 'r02_launches_syn20m.csv')
if os.path.isfile(path):
    tot, n = collections.OrderedDict(), 0
    for row in csv.DictReader([ln for ln in open(path) if ln.startswith('"')]):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        k = re.sub(r'<.*', '', short(row['Kernel Name']))
        d = tot.setdefault(k, [0, 0.0])
        d[0] += 1
        d[1] += ms_of(row['Metric Value'], row['Metric Unit'])
        n += 1
    total = sum(v[1] for v in tot.values())
    OWN = ('spmm_csr', 'gemm_nt', 'decoder_', 'colsum_', 'compact_', 'splitk_', 'select_', 'keep_flags', 'center_normalize',
           'scan_', 'expand_rows', 'pack_rows', 'csr_', 'sort_', 'knn_', 'topk_', 'degree_')
    own = sum(v[1] for k, v in tot.items() if k.startswith(OWN))
    bench = {}
    bp = os.path.join(G, 'r02_bench_syn20m.json')
    if os.path.isfile(bp):
        bench = json.loads(open(bp).read().strip().splitlines()[-1])
        shutil.copy(bp, os.path.join(P, 'r02_bench_syn20m.json'))
    md = ['# Round 2 -- ncu launch list of ONE training step at syn20m (final state of the round)', '',
          'Command (gpurun, 1x B200): `DG_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none '
          '--csv --log-file gpurun_out/r02_launches_syn20m.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eager` (bench.py had exited '
          '0 without ncu directly before). The same step timed with CUDA events without ncu: %s ms (graph replay).'
          % bench.get('ms_per_step', '?'), '',
          '%d launches, %.2f ms summed; own kernels %.1f %% of it. Per-launch times under ncu are cold-cache and serialised: compare SHARES, '
          'not absolutes.' % (n, total, 100 * own / total), '', '| kernel | launches | ms | share |', '|---|---:|---:|---:|']
    for k, (c, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:36]:
        md.append('| `%s` | %d | %.3f | %.1f%% |' % (k, c, ms, 100 * ms / total))
    open(os.path.join(P, 'r02_launches_syn20m_step.md'), 'w').write('\n'.join(md) + '\n')
    shutil.copy(path, os.path.join(P, 'r02_launches_syn20m_step.csv'))
for src, dst in (('r02_bench_full.json', 'r02_bench_syn20m_full_line.json'), ('r02_bench_2gpu.json', 'r02_bench_syn20m_2gpu.json'),
                 ('r02_bench_reference.json', 'r02_bench_reference_arm.json')):
    if os.path.isfile(os.path.join(G, src)):
        lines = [ln for ln in open(os.path.join(G, src)).read().splitlines() if ln.startswith('{')]
        if lines:
            open(os.path.join(P, dst), 'w').write(lines[-1] + '\n')        # the JSON line only (NCCL banner dropped)

# the microbenchmark file is indented JSON: copied whole
if os.path.isfile(os.path.join(G, 'l2_peak.json')):
    json.dump(json.load(open(os.path.join(G, 'l2_peak.json'))), open(os.path.join(P, 'l2_peak.json'), 'w'), indent=1)
