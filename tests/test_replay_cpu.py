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
        assert float(s[2][0]) == m[2] and float(s[2][1]) == m[3]
        assert abs(float(s[3]) - m[4]) < 1e-6
        assert torch.equal(s[4], m[5]) and torch.equal(s[5], m[6]) and float(s[6]) == m[7]
    batch = mem.random_access(20)
    assert batch[0].shape == (20, 3, 3, 2) and batch[2].shape == (20, 2) and batch[6].shape == (20,)
