"""Figures of the simulator and the trainer (simulation_framework.py:218-336, training_framework.py:266-295, 354-377:
FER / BER / average-iteration / simulation-time curves against SNR, training history, gradient statistics).

Host-side presentation only -- nothing here touches the decode path.  matplotlib is imported when a figure is asked for;
without it the call raises ImportError (the result objects and history lists hold the data either way)."""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

# (result field, y label, title, logarithmic y axis)
PANELS = {
    "fer": ("frame_error_rates", "Frame Error Rate (FER)", "Frame Error Rate vs SNR", True),
    "ber": ("bit_error_rates", "Bit Error Rate (BER)", "Bit Error Rate vs SNR", True),
    "iterations": ("average_iterations", "Average Iterations", "Average Iterations vs SNR", False),
    "time": ("simulation_times", "Simulation Time (s)", "Simulation Time vs SNR", False),
}


def _pyplot():
    try:
        import matplotlib.pyplot as plt
    except ImportError as exc:   # pragma: no cover - depends on the environment
        raise ImportError("plotting needs matplotlib, which is not installed; the SimulationResult / history objects "
                          "hold the numbers") from exc
    return plt


def _draw_curves(ax, results: Dict[str, object], panel: str, log_scale: Optional[bool] = None) -> None:
    field, ylabel, title, log_default = PANELS[panel]
    log = log_default if log_scale is None else (log_scale and log_default)
    for name, res in results.items():
        xs, ys = list(res.snr_values), list(getattr(res, field))
        (ax.semilogy if log else ax.plot)(xs[:len(ys)], ys, marker="o", linewidth=2, markersize=5, label=name)
    ax.set_xlabel("SNR (dB)")
    ax.set_ylabel(ylabel)
    ax.set_title(title)
    ax.grid(True, alpha=0.3)
    ax.legend()


def _finish(plt, save_path: Optional[str], dpi: Optional[int] = 300) -> None:
    plt.tight_layout()
    if save_path:
        plt.savefig(save_path, dpi=dpi, bbox_inches="tight")
    plt.show()


def curves(results: Dict[str, object], panel: str, save_path: Optional[str] = None, log_scale: Optional[bool] = None) -> None:
    """One figure with one curve per decoder."""
    plt = _pyplot()
    fig, ax = plt.subplots(figsize=(10, 8))
    _draw_curves(ax, results, panel, log_scale)
    _finish(plt, save_path)


def comparison(results: Dict[str, object], save_path: Optional[str] = None,
               panels: Sequence[str] = ("fer", "ber", "iterations", "time")) -> None:
    """The 2 x 2 overview: FER, BER, average iterations and simulation time against SNR."""
    plt = _pyplot()
    fig, axes = plt.subplots(2, 2, figsize=(15, 12))
    for ax, panel in zip([axes[0][0], axes[0][1], axes[1][0], axes[1][1]], panels):
        _draw_curves(ax, results, panel)
    _finish(plt, save_path)


def series(columns: Sequence[Tuple[str, str, Sequence[float]]], save_path: Optional[str] = None, xlabel: str = "Epoch") -> None:
    """Side-by-side line plots: (title, y label, values) per column (training loss / accuracy / gradient norm)."""
    plt = _pyplot()
    fig, axes = plt.subplots(1, len(columns), figsize=(5 * len(columns), 5))
    axes = [axes] if len(columns) == 1 else list(axes)
    for ax, (title, ylabel, values) in zip(axes, columns):
        ax.plot(list(values))
        ax.set_title(title)
        ax.set_xlabel(xlabel)
        ax.set_ylabel(ylabel)
        ax.grid(True)
    _finish(plt, save_path, dpi=None)


def gradient_analysis(results: Dict[str, Sequence[float]], save_path: Optional[str] = None) -> None:
    """Histogram of the gradient magnitudes and their scatter against the iteration counts."""
    plt = _pyplot()
    fig, axes = plt.subplots(1, 2, figsize=(12, 5))
    axes[0].hist(list(results["gradient_magnitudes"]), bins=20, alpha=0.7)
    axes[0].set_title("Gradient Magnitude Distribution")
    axes[0].set_xlabel("Gradient Magnitude")
    axes[0].set_ylabel("Frequency")
    axes[0].grid(True)
    axes[1].scatter(list(results["iteration_counts"]), list(results["gradient_magnitudes"]), alpha=0.6)
    axes[1].set_title("Gradient Magnitude vs Iterations")
    axes[1].set_xlabel("Iterations")
    axes[1].set_ylabel("Gradient Magnitude")
    axes[1].grid(True)
    _finish(plt, save_path, dpi=None)
