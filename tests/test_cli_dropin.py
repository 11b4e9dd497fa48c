"""The drop-in claim at the CLI level: the reference's UNMODIFIED `main.py --mode minimizer` drives this
package once `src/genome_minimizer_2/minimizer/minimizer_2.py` is replaced by the re-export that
INTEGRATION.md shows (taken verbatim from that file), and writes the FASTA the reference itself wrote
(golden fixtures), with the same progress lines and the reference's exit status (1 even on success,
SURVEY.md F9).

Runs only where the reference tree is mounted (/root/reference: the build container; the GPU box does not
have it — the `-m gpu` tests cover the same entry functions there through the real engine).  There is no
GPU here, so the engine behind the entry functions is a test double backed by the oracle: what is under
test is the wiring — main.py's flags -> entry functions -> loaders (native GenBank scanner, native list
tokeniser) -> plan/drain protocol -> files, stdout, return dict.  CPU only."""
from __future__ import annotations

import os
import re
import shutil
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden

REFERENCE = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REFERENCE, "main.py")),
                                reason="the reference tree is not mounted here")

STUBS = {
    "Bio/__init__.py": "",
    "Bio/SeqIO.py": "def read(*a, **k):\n    raise AssertionError('the drop-in must not call Biopython')\n",
    "Bio/SeqRecord.py": "class SeqRecord:\n    pass\n",
    "matplotlib/__init__.py": "def use(*a, **k):\n    pass\n",
    "matplotlib/pyplot.py": "def ioff(*a, **k):\n    pass\n",
    "seaborn/__init__.py": "",
}


def _shim_from_integration_md() -> str:
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"```python\n(# src/genome_minimizer_2/minimizer/minimizer_2\.py[^\n]*\n.*?)```", text, re.S)
    assert m, "INTEGRATION.md no longer shows the re-export module"
    return m.group(1)


@pytest.fixture(scope="module")
def patched_reference(tmp_path_factory):
    """A writable copy of the reference (its modules mkdir under the tree at import) with the
    one file replaced that INTEGRATION.md says to replace; main.py and everything else untouched."""
    base = tmp_path_factory.mktemp("cli")
    ref = base / "ref"
    shutil.copytree(REFERENCE, ref, ignore=shutil.ignore_patterns(".git", "__pycache__"))
    stubs = base / "stubs"
    for rel, body in STUBS.items():
        p = stubs / rel
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_text(body)
    target = ref / "src" / "genome_minimizer_2" / "minimizer" / "minimizer_2.py"
    assert target.exists()
    target.write_text(_shim_from_integration_md() +
                      "\n# test only: no GPU in this container\nimport engine_double\nengine_double.install()\n")
    return ref, stubs


def _run_cli(patched, args, cwd_files, launcher=()):
    ref, stubs = patched
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(stubs), ROOT, os.path.join(ROOT, "tests")]), MPLBACKEND="Agg")
    return subprocess.run([sys.executable, *launcher, "main.py", "--mode", "minimizer", *args], cwd=ref, env=env,
                          capture_output=True, text=True, timeout=600)


def _inputs(case, d):
    gb = d / "genome.gb"
    gb.write_text(case["genbank"])
    arr = np.empty(len(case["lists"]), dtype=object)
    for i, l in enumerate(case["lists"]):
        arr[i] = l
    npy = d / "genes.npy"
    np.save(npy, arr, allow_pickle=True)
    return str(gb), str(npy)


def _strip_ts(data: str) -> str:
    lines = data.split("\n")
    assert lines[2].startswith("# Generated on: ")
    lines[2] = "# Generated on: <TS>"
    return "\n".join(lines)


@pytest.mark.parametrize("name", ["kat_appB", "kat_stats_quirk", "hundred_and_one"])
def test_unmodified_main_py_single_file(name, patched_reference, tmp_path):
    case = load_golden(name)
    gb, npy = _inputs(case, tmp_path)
    out_dir = tmp_path / "out"
    r = _run_cli(patched_reference, ["--genome-path", gb, "--genes-path", npy, "--single-file",
                                     "--output-dir", str(out_dir), "--model-name", case["model_name"]], tmp_path)
    assert r.returncode == 1, r.stderr[-2000:]                   # F9: main() returns None -> exit status 1
    assert "✗" not in r.stdout, r.stdout[-2000:]
    fasta = out_dir / f"minimized_genomes_{case['model_name']}.fasta"
    assert _strip_ts(fasta.read_text()) == case["single_file"]
    assert case["single_stdout"] in r.stdout                     # the entry function's own lines, in order
    ret = case["single_return"]
    assert f"- Processed: {ret['genome_count']} genomes" in r.stdout
    assert f"- Average percentage reduction: {ret['average_reduction_pct']:.1f}%" in r.stdout
    assert f"- Average genome length: {ret['average_length_bp']:,.1f} bp" in r.stdout


def test_unmodified_main_py_output_file_flag(patched_reference, tmp_path):
    case = load_golden("kat_appB")
    gb, npy = _inputs(case, tmp_path)
    target = tmp_path / "deep" / "er" / "named.fasta"
    r = _run_cli(patched_reference, ["--genome-path", gb, "--genes-path", npy, "--output-file", str(target),
                                     "--model-name", case["model_name"]], tmp_path)
    assert r.returncode == 1 and "✗" not in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    assert _strip_ts(target.read_text()) == case["single_file"]


@pytest.mark.parametrize("name", ["rand_small_2"])
def test_unmodified_main_py_multi_file(name, patched_reference, tmp_path):
    case = load_golden(name)
    gb, npy = _inputs(case, tmp_path)
    out_dir = tmp_path / "many"
    r = _run_cli(patched_reference, ["--genome-path", gb, "--genes-path", npy, "--output-dir", str(out_dir),
                                     "--model-name", case["model_name"]], tmp_path)
    assert r.returncode == 1 and "✗" not in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    got = {fn: (out_dir / fn).read_text() for fn in sorted(os.listdir(out_dir))}
    assert got == case["multi_files"]
    assert case["multi_stdout"].replace("<OUTDIR>", str(out_dir)) in r.stdout


def test_missing_inputs_are_reported_by_main_py(patched_reference, tmp_path):
    r = _run_cli(patched_reference, ["--genome-path", str(tmp_path / "absent.gb"), "--genes-path", "x.npy"], tmp_path)
    assert r.returncode == 1 and "✗ Genome file not found" in r.stdout


def test_unmodified_main_py_under_torchrun_shards_the_samples(patched_reference, tmp_path):
    """`torchrun --nproc-per-node 2 main.py --mode minimizer ...`: the entry function notices the ranks,
    joins the job (gloo here, NCCL on GPUs) and every rank writes its own part of the one file."""
    case = load_golden("hundred_and_one")
    gb, npy = _inputs(case, tmp_path)
    out_dir, log_dir = tmp_path / "out", tmp_path / "logs"
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    launcher = ("-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                "--master-port", str(port), "--redirects", "1", "--log-dir", str(log_dir))
    r = _run_cli(patched_reference, ["--genome-path", gb, "--genes-path", npy, "--single-file",
                                     "--output-dir", str(out_dir), "--model-name", case["model_name"]], tmp_path,
                 launcher=launcher)
    fasta = out_dir / f"minimized_genomes_{case['model_name']}.fasta"
    assert fasta.exists(), r.stdout[-3000:] + r.stderr[-3000:]
    assert _strip_ts(fasta.read_text()) == case["single_file"]
    # per-rank stdout (torchrun --redirects): rank 0 alone prints the entry function's progress lines
    logs = {}
    for root, _dirs, files in os.walk(log_dir):
        if "stdout.log" in files:
            logs[os.path.basename(root)] = open(os.path.join(root, "stdout.log")).read()
    assert sorted(logs) == ["0", "1"], sorted(logs)
    assert logs["0"].count(case["single_stdout"]) == 1
    assert "genes present" not in logs["1"]
    for text in logs.values():                                     # both ranks finish main.py's runner
        assert "✗" not in text and text.count("✓ GENOME MINIMIZATION COMPLETED!") == 1
