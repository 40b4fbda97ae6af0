"""bench.py without a GPU: the reference arm's JSON contract (it runs the oracle port on the host cores) and a static
guard against the one bug class that only shows under torchrun — a local that is `del`eted (or only bound inside a
`world > 1` block) and read later in the same function."""
import ast
import json
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
                        '--cpu-links-per-core', '2'], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'links/s' and d['higher_is_better'] is True and d['value'] > 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
    assert d['e2e'] == dict(value=d['value'], unit='links/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert 'PubMed' in d['config']['workload'] and d['metric'].startswith('precomputed target links/sec')


def test_no_local_is_read_after_being_deleted():
    """`del name` makes later reads of `name` in the same function raise UnboundLocalError; bench.py's multi-GPU block once
    shadowed a flag that way (only ranks launched by torchrun run that block)."""
    for rel in ('bench.py', '__graft_entry__.py', os.path.join('s3grl_b200', 'engine.py'), os.path.join('s3grl_b200', 'loader.py'),
                os.path.join('s3grl_b200', 'dataset.py'), os.path.join('s3grl_b200', 'tuned_sign.py')):
        tree = ast.parse(open(os.path.join(ROOT, rel)).read())
        for fn in [n for n in ast.walk(tree) if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef))]:
            deleted = {}
            for node in ast.walk(fn):
                if isinstance(node, ast.Delete):
                    for t in node.targets:
                        if isinstance(t, ast.Name):
                            deleted.setdefault(t.id, []).append(node.lineno)
            for name, lines in deleted.items():
                last_del = max(lines)
                rebinds = [n.lineno for n in ast.walk(fn) if isinstance(n, ast.Name) and n.id == name
                           and isinstance(n.ctx, ast.Store) and n.lineno > last_del]
                reads = [n.lineno for n in ast.walk(fn) if isinstance(n, ast.Name) and n.id == name
                         and isinstance(n.ctx, ast.Load) and n.lineno > last_del]
                bad = [ln for ln in reads if not any(rb <= ln for rb in rebinds)]
                assert not bad, f"{rel}:{fn.name}: '{name}' deleted at line {last_del} and read at {bad}"
