"""Host-side tools that turn profiler output into the summaries kept under profiles/.  CPU only."""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_launch_summary_reads_the_committed_launch_list():
    csv = ROOT / "profiles" / "r02_launches_bench_steps2_warmup3.csv"
    out = subprocess.run([sys.executable, str(ROOT / "tools" / "launch_summary.py"), str(csv)],
                         capture_output=True, text=True, check=True).stdout
    lines = out.splitlines()
    assert lines[0].startswith("our kernels:")
    top = lines[2].split()
    assert top[-1].startswith("match_top2_kernel") or "match_top2_kernel" in lines[2]
    assert float(top[1].rstrip("%")) > 95.0          # the matcher is the step
    assert any("hough_vote_kernel" in ln for ln in lines) and any("affine_verify" in ln for ln in lines)
