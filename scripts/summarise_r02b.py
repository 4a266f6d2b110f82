"""Final round-2 profiles: the ncu launch lists of one eager step at syn20m and at the lrssl shape (scripts/r02b_capture.sh)
as tracked summaries under profiles/, with the share of kernel time spent in this repo's own kernels (namespace `dg::`)
against library kernels (ATen / cuBLAS / cutlass), and the bench lines of the same state.
    python scripts/summarise_r02b.py
"""
import collections
import csv
import json
import os
import re
import shutil

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(REPO, 'gpurun_out'), os.path.join(REPO, 'profiles')


def ms_of(v, u):
    v = float(v.replace(',', ''))
    return v / 1e6 if u.startswith('n') else v / 1e3 if u.startswith('u') else v * 1e3 if u.startswith('s') else v


def family(name):
    """(own kernel?, short family name)."""
    n = re.sub(r'\s+', ' ', name).replace('void ', '')
    own = n.startswith('dg::')
    if own:
        return True, re.sub(r'[<(].*', '', n.replace('dg::', '').replace('<unnamed>::', ''))
    m = re.search(r'(CUDAFunctor_add|direct_copy|FillFunctor|MulFunctor|normal_and_transform|random_from_to|random_kernel|'
                  r'fused_dropout|CatArray|multi_tensor_apply|reduce_kernel|splitKreduce|epilogue::globalKernel|gemmSN|'
                  r'sgemm|d884gemm|neg_kernel|sigmoid|clamp|reciprocal)', n)
    return False, 'lib: ' + (m.group(1) if m else re.sub(r'[<(].*', '', n)[:40])


def launch_list(tag, title, bench_file):
    path = os.path.join(G, 'r02b_launches_%s.csv' % tag)
    if not os.path.isfile(path):
        return
    tot, n, own_ms, own_n = collections.OrderedDict(), 0, 0.0, 0
    for row in csv.DictReader([ln for ln in open(path) if ln.startswith('"')]):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        own, k = family(row['Kernel Name'])
        ms = ms_of(row['Metric Value'], row['Metric Unit'])
        d = tot.setdefault(k, [0, 0.0, own])
        d[0] += 1
        d[1] += ms
        n += 1
        if own:
            own_ms += ms
            own_n += 1
    total = sum(v[1] for v in tot.values())
    bench = {}
    bp = os.path.join(G, bench_file)
    if os.path.isfile(bp):
        lines = [ln for ln in open(bp).read().splitlines() if ln.startswith('{')]
        bench = json.loads(lines[-1]) if lines else {}
    md = ['# Round 2 (final state) -- ncu launch list of ONE eager training step, %s' % title, '',
          'Command (gpurun, 1x B200): `DG_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none '
          '--csv --log-file gpurun_out/r02b_launches_%s.csv python bench.py --workload %s --steps 1 --warmup 3 --no-cpu-baseline --no-extra '
          '--eager` (scripts/r02b_capture.sh; the same command had exited 0 without ncu directly before). The timed configuration '
          'replays the same step as one CUDA graph: %s ms per step with CUDA events, without ncu (`%s`).'
          % (tag, tag, bench.get('ms_per_step', '?'), bench_file), '',
          '%d launches, %.3f ms summed. **Own kernels (namespace `dg::` of libdreamgnn.so): %d launches, %.1f %% of the kernel time**; '
          'library kernels (ATen elementwise / random / cat, cuBLAS) the rest. Per-launch times under ncu are cold-cache and '
          'serialised: compare SHARES, not absolutes.' % (n, total, own_n, 100 * own_ms / total), '',
          '| kernel | own | launches | ms | share |', '|---|:-:|---:|---:|---:|']
    for k, (c, ms, own) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:48]:
        md.append('| `%s` | %s | %d | %.3f | %.1f%% |' % (k, 'x' if own else '', c, ms, 100 * ms / total))
    open(os.path.join(P, 'r02b_launches_%s_step.md' % tag), 'w').write('\n'.join(md) + '\n')
    shutil.copy(path, os.path.join(P, 'r02b_launches_%s_step.csv' % tag))
    print(tag, n, 'launches', round(total, 3), 'ms, own %.1f%%' % (100 * own_ms / total))


launch_list('syn20m', 'syn20m (100 k x 50 k nodes, 20 M pairs)', 'r02b_bench_full.json')
launch_list('lrssl', 'lrssl shape (763 x 681 nodes, 467 641 pairs)', 'r02b_bench_lrssl.json')
for src, dst in (('r02b_bench_full.json', 'r02b_bench_syn20m_full_line.json'), ('r02b_bench_reference.json', 'r02b_bench_reference_arm.json'),
                 ('r02b_bench_lrssl.json', 'r02b_bench_lrssl.json'), ('r02b_bench_gdataset.json', 'r02b_bench_gdataset.json'),
                 ('r02b_bench_cdataset.json', 'r02b_bench_cdataset.json'), ('r02b_bench_2gpu.json', 'r02b_bench_syn20m_2gpu.json'),
                 ('r02b_bench_8gpu.json', 'r02b_bench_syn20m_8gpu.json')):
    if os.path.isfile(os.path.join(G, src)):
        lines = [ln for ln in open(os.path.join(G, src)).read().splitlines() if ln.startswith('{')]
        if lines:
            open(os.path.join(P, dst), 'w').write(lines[-1] + '\n')
