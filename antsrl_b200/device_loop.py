"""The callers either side of `RLApi.step` kept on the device: action selection of the
reference's agents (`agents/collect_agent.py:161-177`) over a batch of environments, without the observations or the
actions crossing PCIe.

    get_action:  if random.random() > epsilon or not training:  rotation  = argmax(q_rotation)  - rotations // 2
                                                                pheromone = argmax(q_pheromone)
                 else:                                          rotation  = randint(0, rotations) - rotations // 2
                                                                pheromone = randint(0, pheromones)      (per ant)

The reference draws ONE uniform per call, so a whole population either follows the network or explores; with E
independent environments (E copies of main.py's loop) the draw is per environment.  The random branch is
`ants_sample_actions` (Philox keyed by seed, global env id, timestep, ant: independent of the sharding), the per-env
draw comes from a seeded host generator (E values per step: a few bytes, the only host->device traffic of the loop).
PyTorch is plumbing here (argmax, where); the policy network itself is the caller's."""
import numpy as np


class DeviceActionSelector:
    def __init__(self, batch, epsilon, rotations=3, pheromones=3, seed=0):
        import torch
        self._t = torch
        self.batch = batch
        self.epsilon = float(epsilon)
        self.rotations, self.pheromones = int(rotations), int(pheromones)
        self.seed = int(seed)
        self._rs = np.random.RandomState(self.seed & 0x7FFFFFFF)

    def explore_mask(self):
        """(E,) bool on the host: which environments explore this step (collect_agent.py:162: `random() > epsilon`
        follows the network)."""
        return ~(self._rs.random_sample(self.batch.E) > self.epsilon)

    def select(self, q_rotation, q_pheromone, training=True, explore=None):
        """q_rotation (E, N, rotations) / q_pheromone (E, N, pheromones) CUDA tensors (any float dtype), or None for a
        head the policy does not have.  -> (rotation, pheromone) int8 CUDA tensors (E, N) as `BatchedAnts.step` takes
        them, and the host mask of the exploring environments."""
        t = self._t
        E, N = self.batch.E, self.batch.N
        if explore is None:
            explore = self.explore_mask() if training else np.zeros(E, dtype=bool)
        explore = np.asarray(explore, dtype=bool)
        rot = ph = None
        if q_rotation is not None:
            rot = (q_rotation.reshape(E, N, self.rotations).argmax(dim=2) - self.rotations // 2).to(t.int8)
        if q_pheromone is not None:
            ph = q_pheromone.reshape(E, N, self.pheromones).argmax(dim=2).to(t.int8)
        if explore.any():
            r_rot, r_ph = self.batch.sample_actions(self.seed, self.rotations, self.pheromones)
            m = t.from_numpy(explore).to(self.batch.device).reshape(E, 1)
            rot = r_rot.clone() if rot is None else t.where(m, r_rot, rot)
            ph = r_ph.clone() if ph is None else t.where(m, r_ph, ph)
        return rot, ph, explore
