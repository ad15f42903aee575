"""DeviceReplayMemory ring-buffer semantics (agents/replay_memory.py:60-114) on CPU tensors: the same class the GPU
loop uses, against a plain list model of "rolling write, oldest entries overwritten"."""
import numpy as np
import torch

from antsrl_b200.replay import DeviceReplayMemory


def test_ring_buffer_matches_list_model():
    rs = np.random.RandomState(0)
    mem = DeviceReplayMemory(37, (3, 3, 2), (2,), 2, device="cpu")
    model = [None] * 37
    head = 0
    total = 0
    for it in range(9):
        n = int(rs.randint(1, 30))
        st = torch.from_numpy(rs.random_sample((n, 3, 3, 2)).astype(np.float32))
        ag = torch.from_numpy(rs.random_sample((n, 2)).astype(np.float32))
        rot = torch.from_numpy(rs.randint(-1, 2, size=n).astype(np.int8))
        ph = None if it % 4 == 3 else torch.from_numpy(rs.randint(0, 3, size=n).astype(np.int8))
        rew = torch.from_numpy(rs.random_sample(n))
        nst = torch.from_numpy(rs.random_sample((n, 3, 3, 2)).astype(np.float32))
        nag = torch.from_numpy(rs.random_sample((n, 2)).astype(np.float32))
        done = it == 8
        mem.extend(st, ag, (rot, ph), rew, nst, nag, done)
        for k in range(n):
            model[head] = (st[k], ag[k], float(rot[k]), 1.0 if ph is None else float(ph[k]), float(rew[k]), nst[k], nag[k],
                           float(done))
            head = (head + 1) % 37
        total += n
        assert mem.head == head and len(mem) == min(total, 37)
    for slot, m in enumerate(model):
        if m is None:
            continue
        s = mem[slot]
        assert torch.equal(s[0], m[0]) and torch.equal(s[1], m[1])
        assert float(s[2][0]) == m[2] and float(s[2][1]) == m[3] and s[2].dtype == torch.int64
        assert abs(float(s[3]) - m[4]) < 1e-6
        assert torch.equal(s[4], m[5]) and torch.equal(s[5], m[6]) and float(s[6]) == m[7]
    batch = mem.random_access(20)
    assert batch[0].shape == (20, 3, 3, 2) and batch[2].shape == (20, 2) and batch[6].shape == (20,)
    assert batch[6].dtype == torch.bool


def _reference_replay_memory():
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import ref_harness
    if not ref_harness.reference_available():
        return None
    import importlib.util
    root = ref_harness.REFERENCE_ROOT
    if os.path.isfile(root):                                   # oracle/_ref/reference.zip (sourceless .pyc archive)
        import zipimport
        spec = zipimport.zipimporter(os.path.join(root, "agents")).find_spec("replay_memory")
    else:
        spec = importlib.util.spec_from_file_location("ref_replay_memory", os.path.join(root, "agents", "replay_memory.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.ReplayMemory


def test_pinned_against_the_reference_replay_memory():
    """Identical inputs into agents/replay_memory.py (the unmodified reference class) and DeviceReplayMemory: every
    field of the buffer, head and fill agree after each extend, with and without a pheromone action; dtypes are the
    reference's.  Batches that do not cross the end of the buffer (where the reference garbles the remainder, see the
    module docstring of antsrl_b200/replay.py)."""
    import pytest
    Ref = _reference_replay_memory()
    if Ref is None:
        pytest.skip("needs the reference (checkout or oracle/_ref)")
    rs = np.random.RandomState(3)
    ref = Ref(120, (7, 7, 6), [2], [2])
    mem = DeviceReplayMemory(120, (7, 7, 6), (2,), 2, device="cpu")
    for it in range(6):
        n = 20
        st, nst = rs.random_sample((n, 7, 7, 6)), rs.random_sample((n, 7, 7, 6))            # float64 like RLApi's arrays
        ag, nag = rs.random_sample((n, 2)), rs.random_sample((n, 2))
        rot, ph = rs.randint(0, 3, n), (None if it == 2 else rs.randint(0, 3, n))
        rew = rs.random_sample(n) * 5
        done = it == 5
        ref.extend(st.astype(np.float32), ag.astype(np.float32), (rot, ph), rew.astype(np.float32), nst.astype(np.float32),
                   nag.astype(np.float32), done)
        mem.extend(torch.from_numpy(st), torch.from_numpy(ag), (torch.from_numpy(rot), None if ph is None else torch.from_numpy(ph)),
                   torch.from_numpy(rew), torch.from_numpy(nst), torch.from_numpy(nag), done)
        assert mem.head == ref.head and len(mem) == len(ref)
        for a, b in ((mem.states, ref.states), (mem.agent_states, ref.agent_states), (mem.rewards, ref.rewards),
                     (mem.new_states, ref.new_states), (mem.new_agent_states, ref.new_agent_states)):
            assert a.dtype == b.dtype == torch.float32 and torch.equal(a, b)
        assert mem.actions.dtype == ref.actions.dtype and torch.equal(mem.actions, ref.actions)
        assert mem.dones.dtype == ref.dones.dtype == torch.bool and torch.equal(mem.dones, ref.dones)
    got, want = mem[[3, 50, 119]], ref[[3, 50, 119]]
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    s = mem.random_access(64)
    assert len(set(map(tuple, s[0].reshape(64, -1)[:, :3].tolist()))) == 64            # without replacement
    with pytest.raises(ValueError):
        mem.random_access(121)
