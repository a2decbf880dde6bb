"""Put the UNMODIFIED reference (illuin-tech/colpali, `colpali_engine`) under baseline/_ref/ so that it travels to
the GPU box (baseline/_ref is git-ignored, not gpurun-ignored; the repo's history stays free of reference sources).

    python baseline/install_ref.py [--src /root/reference]

1. `pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref <copy>`
   -- fails in this image: the reference's build backend (hatchling + hatch-vcs, pyproject.toml:1-3) is not installed
   and not in /opt/wheelhouse.
2. Fallback = what that wheel would contain: the pure-Python package directory `colpali_engine/`
   (`[tool.hatch.build.targets.wheel] include = ["colpali_engine"]`, pyproject.toml:8-9), copied byte for byte, plus
   the reference's own offline hot-path tests (tests/utils/test_processing_utils.py, tests/loss/test_li_losses.py, tests/loss/test_bi_losses.py)
   under baseline/_ref/ref_tests/ for the replay test (tests/test_reference_replay_gpu.py).
Used by: bench.py --impl reference / cpu_baseline (kind "reference"), tests/test_reference_replay_gpu.py.
"""
import argparse, os, shutil, subprocess, sys, tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
REF_TESTS = ("tests/utils/test_processing_utils.py", "tests/loss/test_li_losses.py", "tests/loss/test_bi_losses.py")


def install(src: str = "/root/reference", verbose: bool = True) -> str:
    if not os.path.isdir(os.path.join(src, "colpali_engine")):
        raise FileNotFoundError(f"{src}/colpali_engine not found")
    how = "pip"
    tmp = tempfile.mkdtemp(prefix="refcopy_")
    try:
        work = os.path.join(tmp, "ref")
        shutil.copytree(src, work, ignore=shutil.ignore_patterns(".git", "__pycache__"))
        if os.path.isdir(DST):
            shutil.rmtree(DST)
        r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                            "--find-links", "/opt/wheelhouse", "--target", DST, work], capture_output=True, text=True)
        if r.returncode != 0 or not os.path.isdir(os.path.join(DST, "colpali_engine")):
            how = "copy (pip failed: " + (r.stderr.strip().splitlines() or ["?"])[-1][:120] + ")"
            os.makedirs(DST, exist_ok=True)
            shutil.copytree(os.path.join(src, "colpali_engine"), os.path.join(DST, "colpali_engine"),
                            ignore=shutil.ignore_patterns("__pycache__"), dirs_exist_ok=True)
        os.makedirs(os.path.join(DST, "ref_tests"), exist_ok=True)
        for t in REF_TESTS:
            shutil.copy(os.path.join(src, t), os.path.join(DST, "ref_tests", os.path.basename(t)))
        with open(os.path.join(DST, "INSTALL_RECORD.txt"), "w") as f:
            f.write(f"source: {src}\nmethod: {how}\n")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    if verbose:
        print(f"reference installed under {DST} via {how}")
    return DST


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    install(ap.parse_args().src)
