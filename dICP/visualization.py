from dicp_b200.visualization import plot_overlay, plot_map   # noqa: F401
