import os
import sys

# The GPU boxes have no network: the replay test imports the real colpali_engine (hence transformers / huggingface_hub);
# nothing in the suite downloads anything, and nothing should wait on a resolver time-out to find that out.
os.environ.setdefault("HF_HUB_OFFLINE", "1")
os.environ.setdefault("TRANSFORMERS_OFFLINE", "1")
os.environ.setdefault("HF_HUB_DISABLE_TELEMETRY", "1")


def _usable_cores() -> int:
    """min(cpu_count, affinity, cgroup quota).  The GPU boxes show 128 logical CPUs but cap the container at 16
    (cpu.max); the CPU oracle (torch einsum / autograd on the host) is ~80x slower when oneDNN oversubscribes them --
    the GPU suite went from 50 s to 12 minutes when four more oracle-backed tests were added."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except AttributeError:
        pass
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:  # cgroup v2
            quota, period = f.read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        try:  # cgroup v1
            with open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us") as f:
                quota = int(f.read())
            with open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as f:
                period = int(f.read())
            if quota > 0 and period > 0:
                n = min(n, max(1, quota // period))
        except (OSError, ValueError):
            pass
    return n


# numpy's BLAS (oracle.maxsim_f64) and every OpenMP runtime size their pools from the 128 visible CPUs unless told
# otherwise; torch.set_num_threads below only covers ATen.  Set before numpy / torch are imported.
for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_k, str(_usable_cores()))

import numpy as np  # noqa: E402
import pytest  # noqa: E402
import torch  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

torch.set_num_threads(_usable_cores())
try:  # pools that were already created before this file was imported (a plugin importing numpy first)
    from threadpoolctl import threadpool_limits

    threadpool_limits(limits=_usable_cores())
except Exception:  # noqa: BLE001 - best effort
    pass


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def from_bits(a: np.ndarray, shape=None) -> torch.Tensor:
    """uint16 bit patterns -> bf16 tensor."""
    t = torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16)
    return t.reshape(shape) if shape is not None else t


def split_rows(flat: torch.Tensor, lens):
    out, o = [], 0
    for n in lens:
        out.append(flat[o : o + int(n)])
        o += int(n)
    return out
