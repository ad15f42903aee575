"""Mirror of the reference's top-level ``utils`` module (utils.py:5-33): only ``AX`` is on the step path."""
import numpy as np

AX = np.newaxis


def perlin_noise_generator(w, h, offset_x, offset_y, scale=22.0, octaves=2, persistence=0.5, lacunarity=2.0):
    raise NotImplementedError("perlin_noise_generator needs the third-party `noise` package (utils.py:12), which "
                              "is not available; SURVEY.md section 8(f) row 1")


def plot_training(reward, loss):   # utils.py:20-33 (plotting, out of scope; kept so `from utils import *` works)
    import matplotlib.pyplot as plt
    fig = plt.figure(figsize=(100, 300))
    ax = fig.add_subplot(211); ax.plot(reward, color='blue')
    ax.set(title="Mean reward per episode", ylabel="Reward", xlabel="Epoch")
    bx = fig.add_subplot(212); bx.plot(loss, color='red')
    bx.set(title="Mean loss per episode", ylabel="Loss", xlabel="Epoch")
    plt.show()
