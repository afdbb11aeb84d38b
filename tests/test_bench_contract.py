"""The JSON line bench.py prints is a contract with the driver.  Checked here on CPU: the committed line of the last
B200 run (profiles/r02_bench_512_b16.json) and a live `--impl reference` line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
             'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'cpu_baseline'}


def _last_json_line(text):
    return json.loads([l for l in text.strip().splitlines() if l.startswith('{')][-1])


def test_recorded_b200_line_has_every_contract_key():
    line = _last_json_line(open(os.path.join(ROOT, 'profiles', 'r02_bench_512_b16.json')).read())
    assert BASE_KEYS | {'gpu_launches', 'clocks', 'roofline'} <= set(line)
    baseline = json.load(open(os.path.join(ROOT, 'BASELINE.json')))
    assert line['unit'] == 'images/s' and line['higher_is_better'] is True and line['scaling'] == 'weak'
    assert line['metric'].split(' (')[0] in baseline['metric']
    assert line['vs_baseline'] is None                      # BASELINE.md holds no published number for this metric
    assert 'workload' in line['config'] and 'model' not in line['config']
    assert line['steps'] >= 1 and line['warmup'] >= 3 and line['gpu_launches'] > 0
    assert abs(line['value'] - line['config']['global_batch'] / line['ms_per_step'] * 1e3) < 1e-2 * line['value']
    e2e = line['e2e']
    assert {'value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step'} <= set(e2e)
    assert e2e['h2d_bytes_per_step'] >= 16 * 512 * 512 * 4 and e2e['d2h_bytes_per_step'] > 0
    assert e2e['value'] != line['value']                    # measured separately, through host buffers
    roof = line['roofline']
    assert {'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'} <= set(roof)
    assert roof['bound'] in ('hbm', 'tensor') and abs(roof['frac'] - roof['achieved'] / roof['peak']) < 1e-3
    # the headline roofline figure is the STEP-level one: algorithmic bytes of an iteration / measured time / peak
    assert roof['scope'] == 'step'
    assert abs(roof['achieved'] - roof['algorithmic_bytes_per_step'] / (line['ms_per_step'] * 1e-3) / 1e9) < 1.0
    for fam in ('family', 'wgrad_family'):
        assert 0 < roof[fam]['frac'] < 1.2 and roof[fam]['launches_per_step'] > 0
    assert roof['best_launch']['frac'] >= roof['family']['frac']
    assert line['rounds']['n'] >= 3 and len(line['rounds']['ms_per_step']) == line['rounds']['n']
    assert line['gpu_eager_baseline']['value'] > 0 and line['gpu_eager_baseline']['kind'] == 'reference'
    clocks = line['clocks']
    assert {'sm_mhz', 'sm_max_mhz', 'reasons'} <= set(clocks)
    assert not {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'} & set(clocks['reasons'])
    cpu = line['cpu_baseline']
    assert {'value', 'unit', 'cores', 'kind', 'sample'} <= set(cpu) and cpu['kind'] in ('reference', 'port')


def test_reference_arm_line_live():
    out = subprocess.run([sys.executable, 'bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '1'],
                         cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-1500:]
    line = _last_json_line(out.stdout)
    assert line['impl'] == 'reference' and BASE_KEYS <= set(line)
    assert line['unit'] == 'images/s' and line['steps'] == 1 and line['value'] > 0
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['d2h_bytes_per_step'] == 0
    assert line['e2e']['value'] == line['value'] == line['cpu_baseline']['value']
    # 'reference' when the unmodified reference modules are present (/root/reference or oracle/_ref), else the port
    assert line['cpu_baseline']['kind'] in ('reference', 'port') and line['cpu_baseline']['cores'] >= 1
