"""Spike-trace accounting of the hot path's outputs: the reference's power proxy
(`audiozen/metric.py:303-340`, PDF eq. 5) evaluated directly on the traces `forward()` returns, on whatever
device they live on, without the per-layer host synchronisations (`print` of device scalars) of the reference.
"""
from __future__ import annotations

import torch


def _width(trace, i):
    w = getattr(trace, "widths", None)
    return w[i] if w is not None else trace[i].size(-1)


def firing_rates(all_layer_outputs):
    """Mean firing rate of every spiking layer of one sequence model: entries 1..-2 of its trace list
    ([x_norm, h1, ..., hL, proj_out], MSF:115-119).  Returns a 1-D tensor on the traces' device.
    Traces of the streaming schedule (`modeling.LazyOutputs`) carry the number of spikes every recurrence kernel
    counted while it ran (popcount of its ballot words), so no fp32 trace is materialised for the accounting."""
    counts = getattr(all_layer_outputs, "spike_counts", None)
    if counts is not None:
        return counts.to(torch.float32) / float(all_layer_outputs.trace_numel)
    return torch.stack([t.gt(0).float().mean() for t in all_layer_outputs[1:-1]])


def compute_synops(fb_all_layer_outputs, sb_all_layer_outputs, shared_weights=True):
    """Synaptic operations per frame per utterance-row (audiozen/metric.py:303-327): for every spiking layer
    rate * fan_in_width * (next_width + own_width), summed over the full-band and all sub-band models; doubled
    when the gate weights are not shared.  One device->host transfer at the end."""
    total = None
    for trace in [fb_all_layer_outputs] + list(sb_all_layer_outputs):
        rates = firing_rates(trace)
        widths = torch.tensor([_width(trace, i) * (_width(trace, i + 1) + _width(trace, i))
                               for i in range(1, len(trace) - 1)], dtype=torch.float32, device=rates.device)
        s = (rates * widths).sum()
        total = s if total is None else total + s
    val = float(total)
    return val if shared_weights else 2.0 * val


def compute_neuronops(fb_all_layer_outputs, sb_all_layer_outputs):
    """Neuron updates per frame (audiozen/metric.py:330-340): the widths of every entry of every trace list."""
    n = sum(_width(fb_all_layer_outputs, i) for i in range(len(fb_all_layer_outputs)))
    for trace in sb_all_layer_outputs:
        n += sum(_width(trace, i) for i in range(len(trace)))
    return float(n)
