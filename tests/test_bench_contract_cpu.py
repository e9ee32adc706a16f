"""bench.py's reference arm prints exactly one JSON line with the contract's keys (CPU, timing stubbed out)."""
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_line_schema(monkeypatch, capfd):
    import bench
    monkeypatch.setattr(bench, 'time_torch_port', lambda nb, steps, warmup, threads: (1.5, 0.25))
    monkeypatch.setattr(bench, 'best_torch_threads', lambda: (8, 0.1))
    monkeypatch.setattr(bench, '_REAL_STDOUT', None)
    monkeypatch.setattr(sys, 'argv', ['bench.py', '--impl', 'reference', '--steps', '2', '--warmup', '3'])
    monkeypatch.delenv('RANK', raising=False)
    captured = {}
    monkeypatch.setattr(bench, 'emit', lambda line: captured.update(line))
    monkeypatch.setattr(bench, 'quiet_stdout', lambda: None)
    bench.main()
    for k in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
              'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert k in captured, k
    assert captured['impl'] == 'reference' and captured['vs_baseline'] is None and captured['dtype'] == 'f32'
    assert captured['e2e']['h2d_bytes_per_step'] == 0 and captured['e2e']['d2h_bytes_per_step'] == 0
    assert set(captured['cpu_baseline']) >= {'value', 'unit', 'cores', 'kind', 'sample'}
    assert 'workload' in captured['config'] and 'model' not in captured['config']
    json.dumps(captured)


def test_non_zero_ranks_of_the_reference_arm_stay_silent(monkeypatch):
    import bench
    monkeypatch.setenv('RANK', '1')
    monkeypatch.setattr(sys, 'argv', ['bench.py', '--impl', 'reference', '--gpus', '2'])
    called = []
    monkeypatch.setattr(bench, 'emit', lambda line: called.append(line))
    monkeypatch.setattr(bench, 'quiet_stdout', lambda: None)
    bench.main()
    assert called == []
