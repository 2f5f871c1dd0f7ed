"""The committed measurement evidence stays readable by the tools that produced it, and bench.py's accounting knows every
entry point the step calls (CPU only: no kernel is launched here)."""
import json
import os
import subprocess
import sys

from conftest import ROOT

PROF = os.path.join(ROOT, 'profiles')


def test_launch_list_parses_and_covers_one_step():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'dev', 'launchlist.py'),
                          os.path.join(PROF, 'r2_launches_c3_x3.csv')], capture_output=True, text=True, check=True).stdout
    head = [l for l in out.splitlines() if l.startswith('# ') and 'kernels, sum' in l][0]
    n, total = int(head.split()[1]), float(head.split()[4])
    assert 20 <= n <= 45 and 800.0 <= total <= 1400.0, head           # one C3 step: ~30 kernels, ~1.07 ms cold
    rows = [l for l in out.splitlines() if l and not l.startswith('#') and ',' in l][1:]
    shares = sum(float(r.rsplit(',', 1)[1]) for r in rows)
    assert abs(shares - 1.0) < 1e-2
    committed = open(os.path.join(PROF, 'r2_launches_c3_x3_summary.csv')).read()
    assert rows[0].split(',')[0] in committed                          # the summary under profiles/ is this capture's


def test_traffic_json_matches_the_ncu_summary(tmp_path):
    dst = str(tmp_path / 'traffic.json')
    subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_traffic.py'),
                    os.path.join(PROF, 'r2_ncu_full_c3_x3.csv'), dst, 'C3'], capture_output=True, text=True, check=True)
    new = json.load(open(dst))['C3']
    old = json.load(open(os.path.join(PROF, 'r2_traffic.json')))['C3']
    for k, v in new.items():
        if not k.startswith('_'):
            assert old[k] == v, k
    assert 'lnst_conv3x3_bf16x3_tc' in new and new['lnst_conv3x3_bf16x3_tc'] > 1e7


def test_bench_accounts_for_every_entry_point_of_the_final_step():
    sys.path.insert(0, ROOT)
    import bench
    d = json.load(open(os.path.join(PROF, 'r2_bench_c3_final.json')))
    table = d['kernel_table_ms_per_step']
    assert abs(sum(table.values()) - 1.25) < 0.35                       # eager re-issue of a ~1.0 ms step
    for name in table:                                                  # aliases are folded: no fused variant appears twice
        assert name not in bench.ALIASES, name
    for fused, base in bench.ALIASES.items():
        assert base in table or base.replace('_tma', '_box') in table or 'max_box' in fused, (fused, base)
    assert d['roofline']['kernel'] == 'lnst_conv3x3_bf16x3_tc' and 0.15 < d['roofline']['frac'] < 0.3
    assert d['clocks']['samples'] >= 1 and not d['clocks']['reasons']
    assert d['e2e']['h2d_bytes_per_step'] > 2.9e7 and d['e2e']['d2h_bytes_per_step'] > 8e6 and d['gpu_launches'] >= 20 * d['steps']
    for k in ('C1', 'C2', 'C4', 'C5'):
        assert d['configs'][k]['value'] > 0
