"""z-slab decomposition of libmadgpu -- NCCL halo exchange, peer-store halo with its IPC import and handshake, agglomeration on
rank 0, FullMultiGrid on slabs -- emulated on the CPU: every rank is a thread driving its own context of the host build of
csrc/madgpu.cu (tests/mad_host/), NCCL is tests/mad_host/fake_nccl.cpp, stream memory operations are release stores and blocking
waits.  Each case runs tests/mad_host/slab_emulation.py in its own process (a protocol dead-lock aborts after a time-out instead of
hanging the suite) and requires the distributed solve to be bit-identical to the single-context solve of the whole volume.

This is where the 8-rank peer-store configuration that hung on real hardware in round 1 is shown to be
correct as a protocol: sequence numbers, waits and signals line up for 8 ranks, two distributed levels and the agglomeration."""
import os
import subprocess
import sys

import pytest

from util import ROOT

SCRIPT = os.path.join(ROOT, "tests", "mad_host", "slab_emulation.py")


def _run(*args, timeout=900):
    env = dict(os.environ, FAKE_CUDA_WAIT_TIMEOUT_S="120", FAKE_NCCL_TIMEOUT_S="240")
    r = subprocess.run([sys.executable, SCRIPT, *args], capture_output=True, text=True, timeout=timeout, env=env)
    assert r.returncode == 0 and "SLAB_EMULATION_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


def test_two_ranks_nccl_halo():
    out = _run("--world", "2", "--peer", "0", "--smoother", "wj", "--nu", "3")
    assert "'agglomeration_level': 1" in out


def test_four_ranks_peer_halo_two_distributed_levels():
    out = _run("--world", "4", "--peer", "1", "--agglomerate-voxels", "1000")
    assert "'agglomeration_level': 2" in out and "'planes_per_rank': [8, 4, 2]" in out


def test_eight_ranks_peer_halo():
    """The configuration of the round-1 hang: 8 ranks, peer stores from the producing kernels, stream-ordered arrival counters."""
    out = _run("--world", "8", "--peer", "1", "--agglomerate-voxels", "1000")
    assert "'planes_per_rank': [8, 4, 2]" in out


def test_temporal_blocking_on_slabs():
    """Fused Gauss-Seidel sweeps (k_coef_gs_tb) on z-slabs: the frozen halo of a pass includes the ghost planes, the last fused sweep
    stores the slab's boundary planes into the neighbours' ghost planes.  The tile grid of a slab differs from the whole volume's,
    so the comparison is on the converged image and the cycle counts."""
    _run("--world", "2", "--peer", "1", "--shape", "32,24,24", "--nu", "3", "--tb", "3")
    _run("--world", "4", "--peer", "0", "--agglomerate-voxels", "1000", "--nu", "3", "--tb", "3")


def test_fmg_on_slabs_with_a_multi_level_agglomerated_hierarchy():
    """agglomerated_fmg: the FMG recursion below the agglomeration level runs as the sub-context's own FullMultiGrid on rank 0."""
    out = _run("--world", "2", "--peer", "1", "--cycle", "fmg", "--shape", "32,24,24")
    assert "'agglomeration_level': 1" in out


def test_fmg_on_four_slabs_nccl_two_distributed_levels():
    _run("--world", "4", "--peer", "0", "--agglomerate-voxels", "1000", "--cycle", "fmg", "--smoother", "wj")


def test_bounded_kernel_wait_eight_ranks():
    """MADGPU_P2P_WAIT=kernel: the arrival counters are awaited by k_halo_wait (bounded) instead of cuStreamWaitValue32.  Weighted
    Jacobi: the same iteration on slabs as in one context -- identical cycle counts and image."""
    _run("--world", "8", "--peer", "1", "--agglomerate-voxels", "1000", "--wait", "kernel", "--smoother", "wj")


def test_a_silent_rank_is_an_error_on_every_rank_not_a_hang():
    """Fault injection (MADGPU_P2P_TEST_DROP_SIGNAL): rank 2 stops signalling.  With the bounded wait its neighbours time out, the
    flag travels with the next residual-norm all-reduce, and every rank returns the same error in the same cycle."""
    env = dict(os.environ, FAKE_CUDA_WAIT_TIMEOUT_S="120", FAKE_NCCL_TIMEOUT_S="240")
    r = subprocess.run([sys.executable, SCRIPT, "--world", "4", "--peer", "1", "--agglomerate-voxels", "1000", "--wait", "kernel", "--drop", "2:37"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "SLAB_EMULATION_TIMEOUT_REPORTED" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    assert "an arrival-counter wait timed out" in r.stdout
