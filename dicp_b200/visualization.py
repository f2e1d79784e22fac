"""Debug plots — mirror of `dICP/visualization.py` (not on any compute path).  matplotlib is
imported lazily so that the package imports without it."""
import torch


def _plt():
    import matplotlib.pyplot as plt
    return plt


def _as_numpy(pc):
    return pc.detach().cpu().numpy() if isinstance(pc, torch.Tensor) else pc


def plot_overlay(pc1, pc2, c1='b', c2='r', file_name="overlay.png"):
    plt = _plt()
    pc1, pc2 = _as_numpy(pc1), _as_numpy(pc2)
    plt.figure()
    for pc, colour in ((pc1, c1), (pc2, c2)):
        plt.scatter(pc[:, 0], pc[:, 1], s=0.5, c=colour)
    plt.savefig(file_name)


def plot_map(points, color='b', map=None):
    plt = _plt()
    pts = _as_numpy(points)
    plt.scatter(pts[:, 0], pts[:, 1], marker='o', color=color)
    plt.xlim(-4, 6)
    plt.ylim(-2, 10)
    plt.show()
