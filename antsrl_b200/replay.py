"""Replay memory that stays in HBM.

Mirror of the reference's `agents/replay_memory.py:6-114` (a ring buffer of tensors with the reference's dtypes:
float32 states / agent states / rewards, integer actions, bool dones; `extend` writes rolling, `random_access`
samples without replacement), fed with the DEVICE tensors the batched step loop returns instead of numpy arrays, so
an on-device policy trains without the (E*N, 7, 7, C) observations ever crossing PCIe.  PyTorch is plumbing here:
the ring-buffer writes are two slice copies per field on the current stream, the sample indices are drawn on the
device (no host round trip per minibatch).

Pinned against the reference class with identical inputs by tests/test_replay_cpu.py.  One deliberate deviation:
when a batch crosses the end of the buffer, the reference's recursive call passes the already stacked (n, 2)
action array where `extend` expects the (rotation, pheromone) pair (replay_memory.py:99-114), so it stores two
garbled entries and drops the rest of the batch; this class wraps around as the docstring of the reference intends.
"""


class DeviceReplayMemory:
    def __init__(self, max_len, observation_space, agent_space, action_space, device="cuda"):
        import torch
        self._t = torch
        self.max_len = int(max_len)
        self.head = 0
        self.fill = 0
        obs = tuple(observation_space)
        ag = tuple(agent_space)
        self.states = torch.zeros((self.max_len,) + obs, dtype=torch.float32, device=device)       # replay_memory.py:18-24
        self.agent_states = torch.zeros((self.max_len,) + ag, dtype=torch.float32, device=device)
        self.actions = torch.zeros((self.max_len, int(action_space)), dtype=torch.int64, device=device)
        self.rewards = torch.zeros((self.max_len,), dtype=torch.float32, device=device)
        self.new_states = torch.zeros((self.max_len,) + obs, dtype=torch.float32, device=device)
        self.new_agent_states = torch.zeros((self.max_len,) + ag, dtype=torch.float32, device=device)
        self.dones = torch.zeros((self.max_len,), dtype=torch.bool, device=device)

    def __len__(self):                                             # replay_memory.py:26-27
        return self.fill

    def __getitem__(self, idx):                                    # replay_memory.py:29-47
        return (self.states[idx], self.agent_states[idx], self.actions[idx], self.rewards[idx], self.new_states[idx],
                self.new_agent_states[idx], self.dones[idx])

    def random_access(self, n):                                    # replay_memory.py:49-58
        if n > len(self):
            raise ValueError("Sample larger than population")          # what random.sample raises in the reference
        idx = self._t.randperm(len(self), device=self.states.device)[:n]   # without replacement, drawn on the device
        return self[idx]

    def extend(self, states, agent_states, actions, rewards, new_states, new_agent_states, done):
        """replay_memory.py:82-114 for device tensors with a flat leading axis (E*N entries): rolling write, old
        entries are overwritten; `actions` = (rotation, pheromone or None) like the reference."""
        t = self._t
        rot, ph = actions
        rot = rot.reshape(-1).to(t.int64)
        ph = t.ones_like(rot) if ph is None else ph.reshape(-1).to(t.int64)          # replay_memory.py:99-102
        act = t.stack((rot, ph), dim=-1)
        n = act.shape[0]
        fields = ((self.states, states.reshape((n,) + self.states.shape[1:])),
                  (self.agent_states, agent_states.reshape((n,) + self.agent_states.shape[1:])),
                  (self.actions, act), (self.rewards, rewards.reshape(-1)),
                  (self.new_states, new_states.reshape((n,) + self.new_states.shape[1:])),
                  (self.new_agent_states, new_agent_states.reshape((n,) + self.new_agent_states.shape[1:])))
        pos = 0
        while pos < n:
            add = min(self.max_len - self.head, n - pos)
            for dst, src in fields:
                dst[self.head:self.head + add].copy_(src[pos:pos + add])
            self.dones[self.head:self.head + add] = bool(done)
            self.fill = max(self.fill, min(self.max_len, self.head + add))
            self.head = (self.head + add) % self.max_len
            pos += add
