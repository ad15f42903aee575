"""environment/RL_api.py of the reference (lines 15-204): the public boundary of the hot path.  ``observation`` and
``step`` keep the reference's signatures and return numpy float64 arrays of the reference's shapes; the work is
done by the CUDA kernels of antsrl_b200 (ants_observe_host / ants_step_host of the C ABI)."""
import numpy as np
from numpy import ndarray
from typing import List, Optional

from utils import AX
from .environment import Environment, EnvObject
from .ants import Ants
from .rewards.reward import Reward

DELTA = 1.1


class RLVisualization(EnvObject):
    def __init__(self, env: Environment, heatmap):
        super().__init__(env)
        self.heatmap = heatmap


class RLApi(EnvObject):
    def __init__(self, reward: Reward, reward_threshold: float, max_speed: float, max_rot_speed: float,
                 carry_speed_reduction: float, backward_speed_reduction: float):
        super().__init__(None)
        self.reward = reward
        self.reward_threshold = reward_threshold
        self.ants = None
        self.original_ants_position = None
        self.perception_radius = 0
        self.perception_mask = None
        self.perceived_objects: List[EnvObject] = []
        self.perception_coords = None
        self.perception_fwd_delta = 0
        self.max_speed = max_speed
        self.max_rot_speed = max_rot_speed
        self.carry_speed_reduction = carry_speed_reduction
        self.backward_speed_reduction = backward_speed_reduction
        # visualisation aid of the reference (RL_api.py:48-50,144-153); not produced by the CUDA path
        self.save_perceptive_field = False
        self.perceptive_field = None

    def visualize_copy(self, newenv: Environment):
        return RLVisualization(newenv, self.reward.visualization())

    def register_ants(self, new_ants: Ants):
        if self.environment is not None:
            self.environment.detach_object(self)
        self.ants = new_ants
        self.environment = new_ants.environment
        self.environment.add_object(self)
        self.perceived_objects = []
        self.original_ants_position = new_ants.xy
        self.reward.setup(self.ants)

    def setup_perception(self, radius: int, objects: List[EnvObject], mask=None, forward_delta=0):
        if self.environment is not None and self.environment._bridge is not None:
            raise RuntimeError("setup_perception must be called before the first observation / step / update")
        self.perception_radius = radius
        self.perception_mask = mask
        self.perceived_objects = objects
        self.perception_fwd_delta = forward_delta
        self.perception_coords = np.dstack([np.arange(-radius, radius + 1)[AX, :].repeat(2 * radius + 1, 0),
                                            np.arange(-radius, radius + 1)[:, AX].repeat(2 * radius + 1, 1)]).astype(float)
        self.perception_coords *= DELTA

    def observation(self):
        """-> (perception (N, 2r+1, 2r+1, C), agent_state (N, 2), state (N, 2 + P)); updates the reward state."""
        perception, agent_state, state, reward = self.environment.device().observe()
        self.reward._rewards = reward
        return perception, agent_state, state

    def step(self, rotation: Optional[ndarray], on_off_pheromones: Optional[ndarray]):
        """-> (perception, agent_state, reward (N,), done)."""
        perception, agent_state, reward, done = self.environment.device().step(rotation, on_off_pheromones)
        self.reward._rewards = reward
        return perception, agent_state, reward, done
