"""environment/environment.py of the reference (lines 5-47): EnvObject base and Environment.  ``update()`` runs the
whole ordered object update as CUDA kernels; host objects are views of device state."""
from typing import List

from ._bridge import DeviceBridge


class EnvObject:
    def __init__(self, environment: 'Environment'):
        self.environment = environment
        if self.environment is not None:
            self.environment.add_object(self)

    def visualize_copy(self, newenv: 'Environment'):
        return EnvObject(newenv)

    def update(self):
        pass

    def update_step(self):
        return 0

    def _pull(self):
        """Refresh host mirrors from the device if the environment already lives there."""
        env = self.environment
        if env is not None and env._bridge is not None:
            env._bridge.pull()


class Environment:
    def __init__(self, w, h, max_time):
        self.w = w
        self.h = h
        self.objects: List[EnvObject] = []
        self.max_time = max_time
        self._timestep = 1
        self._bridge = None
        # collision noise of Walls.update (walls.py:28): None = draw from the global numpy RNG exactly like the
        # reference (one value per colliding ant, in ant order); "philox" = in-kernel counter-based noise;
        # or a callable(hit_mask) -> values for the colliding ants.
        self.collision_noise = None

    @property
    def timestep(self):
        return self._timestep

    @timestep.setter
    def timestep(self, v):
        self._timestep = int(v)

    def add_object(self, obj: EnvObject):
        self.objects.append(obj)

    def detach_object(self, obj: EnvObject):
        if obj in self.objects:
            self.objects.remove(obj)

    def save_state(self):
        newenv = Environment(self.w, self.h, self.max_time)
        for obj in self.objects:
            newenv.add_object(obj.visualize_copy(newenv))
        return newenv

    def device(self):
        """The CUDA backend of this environment (created on first use from the attached objects)."""
        if self._bridge is None:
            self._bridge = DeviceBridge(self)
        return self._bridge

    def update(self):
        self.device().update()
        self._timestep += 1

    def push_state(self):
        """Upload host-side edits of the objects' arrays to the device."""
        if self._bridge is not None:
            self._bridge.push()
