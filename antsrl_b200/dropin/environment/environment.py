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

    # Pickling (main.py:139-147 pickles the list of save_state() copies; Walls.visualize_copy returns the live
    # object, walls.py:16-17, so the live environment hangs off every snapshot): the pickled state uses the
    # reference's attribute names, so a file written here loads under the reference's classes (the pygame viewer)
    # and a file written by the reference loads here.
    _MIRRORS = ()      # attributes mirrored from the device, kept as _<name> on the host object

    def __getstate__(self):
        self._pull()
        d = dict(self.__dict__)
        for name in self._MIRRORS:
            if "_" + name in d:
                d[name] = d.pop("_" + name)
        return d

    def __setstate__(self, d):
        d = dict(d)
        for name in self._MIRRORS:
            if name in d:
                d["_" + name] = d.pop(name)
        self.__dict__.update(d)


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

    def __getstate__(self):
        """The reference's attributes (environment.py:22-27); the device handle stays behind, the host mirrors of
        the objects are refreshed first, so an unpickled Environment continues on a new handle."""
        if self._bridge is not None:
            self._bridge.pull()
        return {"w": self.w, "h": self.h, "objects": self.objects, "max_time": self.max_time,
                "timestep": self._timestep}

    def __setstate__(self, d):
        d = dict(d)
        self._timestep = int(d.pop("timestep", 1))
        self._bridge = None
        self.collision_noise = None
        self.__dict__.update(d)

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
