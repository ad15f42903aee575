"""Mirror of the reference's top-level ``utils`` module (utils.py:5-33): only ``AX`` is on the step path;
``perlin_noise_generator`` restates the absent third-party ``noise.pnoise2`` (antsrl_b200/perlin.py)."""
import numpy as np

AX = np.newaxis


from antsrl_b200.perlin import perlin_noise_generator   # noqa: E402,F401  (utils.py:7-18; restated, parity unpinned)


def plot_training(reward, loss):   # utils.py:20-33 (plotting, out of scope; kept so `from utils import *` works)
    import matplotlib.pyplot as plt
    fig = plt.figure(figsize=(100, 300))
    ax = fig.add_subplot(211); ax.plot(reward, color='blue')
    ax.set(title="Mean reward per episode", ylabel="Reward", xlabel="Epoch")
    bx = fig.add_subplot(212); bx.plot(loss, color='red')
    bx.set(title="Mean loss per episode", ylabel="Loss", xlabel="Epoch")
    plt.show()
