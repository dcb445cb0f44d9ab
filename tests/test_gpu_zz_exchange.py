"""GPU (two or more devices): the path's one exchange inside the library — rtjx_comm_init / rtjx_gather (exchange.cc, NCCL).
Skipped on single-GPU boxes; the world_size-2 gloo test of the Python plumbing (tests/test_distributed_gloo.py) runs on CPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gather_is_a_no_op_without_a_communicator(golden_dir):
    import numpy as np
    import regtools_b200 as rt
    bam = os.path.join(golden_dir, "kat", "synth.bam")
    ex = rt.JunctionsExtractor(bam, ".", 0)
    ex.identify_junctions_from_BAM()
    want = ex.junction_table()
    ex.gather(0)
    assert np.array_equal(ex.junction_table(), want) and len(want) > 100
    ex.close()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_extract_gathers_to_the_single_gpu_table(world, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    bam = str(tmp_path / "gen.bam")
    subprocess.check_call([os.path.join(ROOT, "tools", "bamgen"), "gen", "--out", bam, "--config", "c3", "--reads", "1500000", "--seed", "21"],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(29611 + world), os.path.join(ROOT, "tests", "exchange_worker.py"), bam],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "EXCHANGE_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]
