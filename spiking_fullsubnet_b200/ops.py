"""Tensor-level wrappers over the C ABI (include/gsn_b200.h).  PyTorch is used only for device
memory and streams: every function takes CUDA fp32 tensors, allocates outputs with torch, and
enqueues the library's kernels on torch's current stream."""
from __future__ import annotations

import torch

from . import _lib

_ACT = {None: 0, False: 0, "tanh": 1, "sigmoid": 2, "relu": 3}
LAUNCHES = [0]   # number of libgsn_b200 kernels enqueued so far (bench.py reports it)
LAST_WS = [None]  # workspace of the last recurrence call (tcgen05 backend: 8 cycle counters of CTA 0)
PROFILE = None   # bench.py sets a list: (algorithmic flops, start event, stop event, (T, R, H)) per recurrence call


def _prep(*tensors):
    """Validate device/dtype/contiguity; bind the library to the tensors' device; return stream."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("spiking_fullsubnet_b200 ops need CUDA tensors (there is no CPU path)")
        if t.dtype != torch.float32:
            raise TypeError(f"expected float32, got {t.dtype}")
        if not t.is_contiguous():
            raise ValueError("expected a contiguous tensor")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise ValueError("tensors on different devices")
    lib = _lib.load()
    _bind(dev)
    return lib, torch.cuda.current_stream(dev).cuda_stream


def _bind(dev):
    """The library launches on the calling thread's CURRENT CUDA device (SM count, tensor-memory planning, stream
    handles belong to it).  No cache: the current device is per-thread state that torch.cuda.set_device / device
    context managers change behind our back, so it is compared on every call and re-bound when it differs."""
    if torch.cuda.current_device() != dev.index:
        _lib.check(_lib.load().gsn_bind_device(dev.index))


OPT_PDL, OPT_F32_MAX_CTAS = _lib.OPT_PDL, _lib.OPT_F32_MAX_CTAS


def set_option(option, value):
    """gsn_set_option for the calling thread (e.g. OPT_PDL around the recurrence launches of a frame-chunk chain)."""
    _lib.check(_lib.load().gsn_set_option(int(option), int(value)))


def _ptr(t):
    return None if t is None else t.data_ptr()


def _spec_layout(z, what):
    """0: contiguous [.., F, T]; 1: time-major (a transposed view of contiguous [.., T, F], which is what torch.stft
    returns and what cuFFT reads and writes)."""
    if z.dtype != torch.complex64 or not z.is_cuda:
        raise ValueError(f"{what}: expected a CUDA complex64 tensor")
    if z.is_contiguous():
        return 0
    if z.transpose(-1, -2).is_contiguous():
        return 1
    raise ValueError(f"{what}: spectrum must be contiguous as [.., F, T] or as [.., T, F]")


def compress_mag(mag, f_keep, fdrc, out=None):
    """mag [B,F,T] (or the complex STFT itself, in either layout of `_spec_layout`) -> cm [T,B,f_keep] = |.|**fdrc
    (MSF:434-436, time-major).  `out`: write into this [T,B,f_keep] buffer."""
    pre = getattr(mag, "_gsn_cm", None)  # gsn_stft_compress already produced it (modeling._stft_fused)
    if pre is not None and pre[0] == int(f_keep) and pre[1] == float(fdrc):
        if out is None:
            return pre[2]
        out.copy_(pre[2])
        return out
    B, F, T = mag.shape
    hook = getattr(mag, "_gsn_cm_capture", None)
    if hook is not None:
        # modeling._GraphedNetwork.network is capturing its graph on this (static) tensor: the compression stays OUTSIDE
        # the graph -- it runs eagerly on the caller's tensor in front of every replay, into the buffer handed out here,
        # so the caller's input is never copied into a static buffer first
        args = (int(f_keep), float(fdrc))
        if hook["args"] is None:
            hook["args"] = args
            hook["cm"] = torch.empty((T, B, f_keep), device=mag.device, dtype=torch.float32)
        if hook["args"] == args:
            return hook["cm"]
        hook["bad"] = True  # a second, different compression of the same input: not a schedule this shortcut knows
    if out is not None and (out.shape != (T, B, f_keep) or out.dtype != torch.float32 or not out.is_contiguous()
                            or out.device != mag.device):
        raise ValueError("compress_mag: out must be a contiguous float32 [T,B,f_keep] tensor on the input's device")
    if mag.is_complex():
        tm = _spec_layout(mag, "compress_mag")
        lib = _lib.load()
        _bind(mag.device)
        st = torch.cuda.current_stream(mag.device).cuda_stream
        cm = out if out is not None else torch.empty((T, B, f_keep), device=mag.device, dtype=torch.float32)
        _lib.check(lib.gsn_compress_spec(mag.data_ptr(), _ptr(cm), B, F, f_keep, T, float(fdrc), tm, st))
        LAUNCHES[0] += 1
        return cm
    lib, st = _prep(mag)
    cm = out if out is not None else torch.empty((T, B, f_keep), device=mag.device, dtype=torch.float32)
    _lib.check(lib.gsn_compress_mag(_ptr(mag), _ptr(cm), B, F, f_keep, T, float(fdrc), st))
    LAUNCHES[0] += 1
    return cm


def _mag_like(mag, out, what):
    """mag_out of the spectral back end: fp32, same shape and (element) strides as the complex `out`."""
    if mag is None:
        return None
    if mag.dtype != torch.float32 or mag.shape != out.shape or mag.stride() != out.stride() or mag.device != out.device:
        raise ValueError(f"{what}: mag must be float32 with the shape and strides of out")
    return mag.data_ptr()


def deepfilter_spec(proj, spec, out, N, ctr, df, S, lo, layout=0, mag=None):
    """Deep filter of one band (MSF:315-346) from its proj output, complex in / complex out: spec [B,F,T] complex64,
    out [B,S,F_out,T] complex64 (bins [lo, lo + N*ctr) are written), both in the same layout of `_spec_layout`.
    layout 0: proj features (c fc df s), MSF:160-167; 1: (c df s fc), cirm_gsn CGN:230."""
    lib, st = _prep(proj)
    tm = _spec_layout(spec, "deepfilter_spec")
    if _spec_layout(out, "deepfilter_spec") != tm or out.device != proj.device or spec.device != proj.device:
        raise ValueError("deepfilter_spec: spec and out must share layout and device")
    T = proj.shape[0]
    B, F, _ = spec.shape
    _lib.check(lib.gsn_deepfilter_spec(_ptr(proj), spec.data_ptr(), out.data_ptr(), _mag_like(mag, out, "deepfilter_spec"),
                                       T, B, N, ctr, df, S, lo, F, out.shape[2], int(layout), tm, st))
    LAUNCHES[0] += 1


def spec_passthrough(spec, out, f_lo, mag=None):
    """out[b, s, f, :] = spec[b, f, :] for f >= f_lo (the bins no band filters, MSF:461-468)."""
    tm = _spec_layout(spec, "spec_passthrough")
    if _spec_layout(out, "spec_passthrough") != tm:
        raise ValueError("spec_passthrough: spec and out must share their layout")
    lib = _lib.load()
    _bind(spec.device)
    st = torch.cuda.current_stream(spec.device).cuda_stream
    B, F, T = spec.shape
    _lib.check(lib.gsn_spec_passthrough(spec.data_ptr(), out.data_ptr(), _mag_like(mag, out, "spec_passthrough"), T, B,
                                        out.shape[1], int(f_lo), F, out.shape[2], tm, st))
    LAUNCHES[0] += 1


def frame_signal(y, window, hop):
    """Analysis half of torch.stft(center=True, pad_mode="constant"): y [B,L] -> windowed frames [B,T,n_fft]."""
    lib, st = _prep(y, window)
    B, L = y.shape
    n_fft = window.numel()
    T = 1 + L // int(hop)
    frames = torch.empty((B, T, n_fft), device=y.device, dtype=torch.float32)
    _lib.check(lib.gsn_frame_signal(_ptr(y), _ptr(window), _ptr(frames), B, L, T, n_fft, int(hop), st))
    LAUNCHES[0] += 1
    return frames


def overlap_add(frames, window, hop, length):
    """Synthesis half of torch.istft(center=True): frames [B,T,n_fft] (irfft of every frame) -> y [B,length]."""
    lib, st = _prep(frames, window)
    B, T, n_fft = frames.shape
    y = torch.empty((B, length), device=frames.device, dtype=torch.float32)
    _lib.check(lib.gsn_overlap_add(_ptr(frames), _ptr(window), _ptr(y), B, T, n_fft, int(hop), int(length), st))
    LAUNCHES[0] += 1
    return y


FFT_FUSED_N = 512  # the transform length gsn_fft.cu is built for (every recipe's n_fft)


def stft_compress(y, window, hop, f_keep=None, fdrc=0.5):
    """torch.stft(center=True, pad_mode="constant") of y [B,L] with the 512-sample `window` AND the network's
    compressed magnitude in one kernel: returns (spec, cm): spec complex64 [B,F,T] as a time-major VIEW of the [B,T,F]
    buffer the kernel writes (the layout `_spec_layout` calls 1), cm [T,B,f_keep] = |spec|**fdrc or None."""
    import ctypes as C
    lib, st = _prep(y, window)
    B, L = y.shape
    n_fft = window.numel()
    T = 1 + L // int(hop)
    F = n_fft // 2 + 1
    spec = torch.empty((B, T, F), device=y.device, dtype=torch.complex64)
    cm = torch.empty((T, B, f_keep), device=y.device, dtype=torch.float32) if f_keep else None
    _lib.check(lib.gsn_stft_compress(_ptr(y), _ptr(window), spec.data_ptr(), _ptr(cm) if cm is not None else None, B, L, T,
                                     n_fft, int(hop), int(f_keep or 0), float(fdrc), st))
    LAUNCHES[0] += 1
    return spec.transpose(1, 2), cm


def irfft_frames(spec):
    """spec complex64 [B,F,T] in the time-major layout -> frames [B,T,512] = torch.fft.irfft of every frame."""
    if _spec_layout(spec, "irfft_frames") != 1:
        raise ValueError("irfft_frames: the spectrum must be time-major ([B,T,F] in memory)")
    lib = _lib.load()
    _bind(spec.device)
    st = torch.cuda.current_stream(spec.device).cuda_stream
    B, F, T = spec.shape
    n_fft = 2 * (F - 1)
    frames = torch.empty((B, T, n_fft), device=spec.device, dtype=torch.float32)
    _lib.check(lib.gsn_irfft_frames(spec.data_ptr(), _ptr(frames), B, T, n_fft, st))
    LAUNCHES[0] += 1
    return frames


def deepfilter_irfft(projs, spec, Ns, ctrs, dfs, layout=0, want_mag=True, want_enh=False):
    """Deep filter of all bands (MSF:315-346, 449-472; one speaker) + pass-through of the bins above them (MSF:461-468)
    + inverse real FFT of every frame in ONE kernel: projs[i] [T, B*Ns[i], 2*ctrs[i]*dfs[i]] in frequency order, spec
    complex64 [B,F,T] time-major.  Returns (frames [B,T,512], mag [B,1,F,T] view or None, enh [B,1,F,T] view or None)."""
    import ctypes as C
    if _spec_layout(spec, "deepfilter_irfft") != 1:
        raise ValueError("deepfilter_irfft: the spectrum must be time-major ([B,T,F] in memory)")
    lib, st = _prep(*projs)
    if spec.device != projs[0].device:
        raise ValueError("deepfilter_irfft: spec and proj on different devices")
    B, F, T = spec.shape
    n_fft = 2 * (F - 1)
    nb = len(projs)
    for p, n, c, d in zip(projs, Ns, ctrs, dfs):
        if p.dtype != torch.float32 or not p.is_contiguous() or p.numel() != T * B * n * 2 * c * d:
            raise ValueError(f"deepfilter_irfft: proj {tuple(p.shape)} is not [T={T}, B*N={B * n}, 2*ctr*df={2 * c * d}]")
    frames = torch.empty((B, T, n_fft), device=spec.device, dtype=torch.float32)
    mag = torch.empty((B, 1, T, F), device=spec.device, dtype=torch.float32) if want_mag else None
    enh = torch.empty((B, 1, T, F), device=spec.device, dtype=torch.complex64) if want_enh else None
    parr = (C.c_void_p * nb)(*[p.data_ptr() for p in projs])
    iarr = [(C.c_int * nb)(*[int(v) for v in vs]) for vs in (Ns, ctrs, dfs)]
    _lib.check(lib.gsn_deepfilter_irfft(parr, iarr[0], iarr[1], iarr[2], nb, int(layout), spec.data_ptr(), _ptr(frames),
                                        _ptr(mag) if mag is not None else None,
                                        enh.data_ptr() if enh is not None else None, B, T, n_fft, st))
    LAUNCHES[0] += 1
    return (frames, mag.transpose(2, 3) if mag is not None else None,
            enh.transpose(2, 3) if enh is not None else None)


def _out(out, shape, like):
    if out is None:
        return torch.empty(shape, device=like.device, dtype=torch.float32)
    if tuple(out.shape) != tuple(shape):
        raise ValueError(f"out has shape {tuple(out.shape)}, expected {tuple(shape)}")
    return out


def subband_features(cm, fb, N, lo, ctr, nbr, ln_weight=None, ln_bias=None, eps=1e-5, out=None):
    """Gather (+reflect, + tiled full-band output) + LayerNorm -> x [T, B*N, K] (MSF:241-312, 111-112)."""
    lib, st = _prep(cm, fb, ln_weight, ln_bias, out)
    T, B, f_cm = cm.shape
    K = ctr + 2 * nbr + (ctr if fb is not None else 0)
    f_fb = fb.shape[2] if fb is not None else 0
    x = _out(out, (T, B * N, K), cm)
    _lib.check(lib.gsn_subband_features(_ptr(cm), f_cm, _ptr(fb), f_fb, _ptr(x), T, B, N, lo, ctr, nbr,
                                        _ptr(ln_weight), _ptr(ln_bias), float(eps), st))
    LAUNCHES[0] += 1
    return x


def subband_rowsums(cm, fb, N, lo, ctr, nbr):
    """Row sums [T, B*N] of the gathered (un-normalised) features `subband_features` would write, without writing them."""
    lib, st = _prep(cm, fb)
    T, B, f_cm = cm.shape
    f_fb = fb.shape[2] if fb is not None else 0
    rs = torch.empty((T, B * N), device=cm.device, dtype=torch.float32)
    _lib.check(lib.gsn_subband_rowsums(_ptr(cm), f_cm, _ptr(fb), f_fb, _ptr(rs), T, B, N, lo, ctr, nbr, st))
    LAUNCHES[0] += 1
    return rs


TC_TRAIN = [True]   # training forward on tcgen05 where supported (False: fp32 CUDA-core kernel)
TC_LINEAR = [True]  # spike-input linears on tcgen05 (set False to force the fp32 CUDA-core kernel)


SPIKE_BITS = [True]  # recurrences also emit the bit-packed trace and spike-input linears read it (False: fp32 trace)


def spike_bits_buffer(shape, H, device):
    """Bit-packed trace buffer for a [..., H] spike trace: int32 [..., ceil(H/32)] (see gsn_layer_recurrence_bits)."""
    return torch.empty(tuple(shape) + ((H + 31) // 32,), device=device, dtype=torch.int32)


def pack_spikes(h):
    """fp32 {0,1} trace [..., H] -> bit-packed int32 [..., ceil(H/32)] (gsn_pack_spikes)."""
    lib, st = _prep(h)
    H = h.shape[-1]
    bits = spike_bits_buffer(h.shape[:-1], H, h.device)
    _lib.check(lib.gsn_pack_spikes(_ptr(h), bits.data_ptr(), h.numel() // H, H, st))
    LAUNCHES[0] += 1
    return bits


def linear(a, w, bias=None, act=None, out=None, out_act=None, spikes=False, sm_budget=0, bits=None):
    """out[..., N] = a[..., K] @ w[N,K]^T + bias.  Returns out, or (out, act(out)) if act.
    spikes=True promises that `a` holds {0,1} (a spike trace): the product then runs on tcgen05 with the
    weights as exact bf16x3 planes (gsn_linear_spikes); otherwise fp32 FMA on CUDA cores (gsn_linear_f32).
    bits: the same trace bit-packed (int32 [..., ceil(K/32)], from layer_recurrence(out_bits=...)); when given
    the tcgen05 kernel reads it instead of `a` (gsn_linear_spike_bits)."""
    lib, st = _prep(a, w, bias, out, out_act)
    K = a.shape[-1]
    N = w.shape[0]
    if w.shape[1] != K:
        raise ValueError(f"linear: a[..., {K}] vs w{tuple(w.shape)}")
    M = a.numel() // K
    out = _out(out, a.shape[:-1] + (N,), a)
    code = _ACT[act]
    out_act = _out(out_act, out.shape, a) if code else None
    if bits is not None and TC_LINEAR[0] and K <= 320:
        if bits.dtype != torch.int32 or not bits.is_contiguous() or bits.numel() != M * ((K + 31) // 32):
            raise ValueError(f"linear: bits {tuple(bits.shape)} {bits.dtype} does not match a{tuple(a.shape)}")
        _lib.check(lib.gsn_linear_spike_bits(bits.data_ptr(), _ptr(w), _ptr(bias), _ptr(out), _ptr(out_act), code,
                                             M, K, N, int(sm_budget), st))
    elif spikes and TC_LINEAR[0] and K % 4 == 0 and K <= 320 and a.data_ptr() % 16 == 0:
        _lib.check(lib.gsn_linear_spikes(_ptr(a), _ptr(w), _ptr(bias), _ptr(out), _ptr(out_act), code, M, K, N,
                                         int(sm_budget), st))
    else:
        _lib.check(lib.gsn_linear_f32(_ptr(a), _ptr(w), _ptr(bias), _ptr(out), _ptr(out_act), code, M, K, N, st))
    LAUNCHES[0] += 1
    return (out, out_act) if code else out


def recurrence_workspace(R, H, shared, backend, device):
    """A 256-byte aligned workspace tensor view for layer_recurrence (reusable across calls of one layer)."""
    nbytes = _lib.load().gsn_layer_recurrence_workspace_bytes(R, H, int(shared), _lib.BACKENDS[backend])
    ws = torch.empty((max(nbytes, 256) + 255) // 4 + 1, device=device, dtype=torch.float32)
    off = ((-ws.data_ptr()) % 256) // 4
    return ws[off:]


def layer_recurrence(xproj, w_hh, bias, bn_scale=None, bn_shift=None, shared=True, want_c=False,
                     h0=None, c0=None, want_state=False, backend="auto", out_h=None, out_c=None, out_hT=None,
                     out_cT=None, workspace=None, sm_budget=0, out_bits=None):
    """One GSULayer over all frames (ESN:75-81 / 132-153).  xproj [T,R,gH] -> h [T,R,H] (and c, (hT,cT)).
    out_bits: optional int32 [T,R,ceil(H/32)] buffer that receives the same spike trace bit-packed."""
    lib, st = _prep(xproj, w_hh, bias, bn_scale, bn_shift, h0, c0, out_h, out_c, out_hT, out_cT)
    T, R, gH = xproj.shape
    H = w_hh.shape[1]
    if gH != (H if shared else 2 * H) or w_hh.shape[0] != gH or bias.numel() != 2 * H:
        raise ValueError(f"layer_recurrence: inconsistent shapes xproj{tuple(xproj.shape)} "
                         f"w_hh{tuple(w_hh.shape)} bias{tuple(bias.shape)} shared={shared}")
    be = _lib.BACKENDS[backend]
    dev = xproj.device
    h = _out(out_h, (T, R, H), xproj)
    c = _out(out_c, (T, R, H), xproj) if (want_c or out_c is not None) else None
    hT = _out(out_hT, (R, H), xproj) if (want_state or out_hT is not None) else None
    cT = _out(out_cT, (R, H), xproj) if (want_state or out_cT is not None) else None
    ws = workspace if workspace is not None else recurrence_workspace(R, H, shared, backend, dev)
    off = 0
    if ws.data_ptr() % 256:
        raise ValueError("workspace must be 256-byte aligned (use ops.recurrence_workspace)")
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    if out_bits is not None and (out_bits.dtype != torch.int32 or not out_bits.is_contiguous()
                                 or tuple(out_bits.shape) != (T, R, (H + 31) // 32)):
        raise ValueError(f"layer_recurrence: out_bits {tuple(out_bits.shape)} {out_bits.dtype}")
    _lib.check(lib.gsn_layer_recurrence_bits(_ptr(xproj), _ptr(w_hh), _ptr(bias), _ptr(bn_scale), _ptr(bn_shift),
                                             _ptr(h0), _ptr(c0), _ptr(h), _ptr(c), _ptr(hT), _ptr(cT),
                                             _ptr(out_bits), T, R, H, int(shared), be, int(sm_budget),
                                             ws.data_ptr() + off, st))
    # kernels enqueued: tcgen05 = the recurrence kernel alone (it also writes the packed trace); SIMT = weight
    # transpose + recurrence; SIMT / int8 pack the trace with one more launch when it is requested
    resolved = be if be != _lib.BACKEND_AUTO else lib.gsn_layer_recurrence_pick_backend(R, H, int(shared))
    LAUNCHES[0] += (2 if resolved == _lib.BACKEND_SIMT else 1) + \
        (1 if out_bits is not None and resolved != _lib.BACKEND_TCGEN05 else 0)
    LAST_WS[0] = (ws, 0)
    if PROFILE is not None:
        e1.record()
        PROFILE.append((2.0 * T * R * gH * H, e0, e1, (T, R, H)))
    return h, c, (hT, cT)


def recurrence_tile(R, H, shared, backend="auto", sm_budget=0):
    """Rows per cluster (16 / 32 / 64) the tcgen05 back ends will use for this shape and SM budget (0: SIMT)."""
    return _lib.load().gsn_layer_recurrence_tile(R, H, int(shared), _lib.BACKENDS[backend], int(sm_budget))


def unpack_spikes(bits, H):
    """bit-packed trace int32 [..., ceil(H/32)] -> fp32 {0,1} [..., H] (inverse of pack_spikes; torch ops, used only
    when a caller asks for the reference-shaped fp32 trace of a streamed layer)."""
    sh = torch.arange(32, device=bits.device, dtype=torch.int32)
    b = (bits.unsqueeze(-1) >> sh) & 1
    return b.reshape(bits.shape[:-1] + (bits.shape[-1] * 32,))[..., :H].to(torch.float32)


def frame_counters(T, device, n=1):
    """n zeroed per-frame counter arrays [n, T] (uint32 as int32) for the streaming pipeline."""
    return torch.zeros((n, T), device=device, dtype=torch.int32)


_PRELOADED = set()


def stream_preload(device):
    """Load the streaming pipeline's kernels on `device` (once): see gsn_stream_preload."""
    idx = torch.device(device).index
    idx = torch.cuda.current_device() if idx is None else idx
    if idx not in _PRELOADED:
        with torch.cuda.device(idx):
            _lib.check(_lib.load().gsn_stream_preload())
        _PRELOADED.add(idx)


def stream_tile(R, H, K_in=0, fused=False, sm_budget=0):
    """Row tile (16 / 32 / 64; 0: unsupported) a gsn_recurrence_stream launch of this shape uses."""
    return _lib.load().gsn_recurrence_stream_tile(R, H, int(K_in), int(bool(fused)), int(sm_budget))


def stream_ctas(R, H, K_in=0, fused=False, sm_budget=0):
    """CTAs of a gsn_recurrence_stream launch = value its out counters reach when a frame is complete (0: unsupported)."""
    return _lib.load().gsn_recurrence_stream_ctas(R, H, int(K_in), int(bool(fused)), int(sm_budget))


def recurrence_stream(w_hh, bias, bn_scale=None, bn_shift=None, xproj=None, in_bits=None, w_ih=None, out_bits=None,
                      out_h=None, out_c=None, out_hT=None, out_cT=None, in_cnt=None, in_target=0, out_cnt=None,
                      spike_count=None, sm_budget=0, workspace=None, in_planes=None, frames_rows=None, planes_ring=0,
                      in_image=None, img_out=None, img_ring=0, bp_cnt=None, bp_target=0):
    """One GSULayer over all frames as a persistent streaming launch (gsn_recurrence_stream): zero initial state,
    shared gate weights.  Input: xproj [T,R,H], OR (in_bits [T,R,ceil(K/32)] int32, w_ih [H,K]) for the fused
    spike-input product, OR (in_planes = the operand images of `xplanes_stream`, w_ih [H,K], frames_rows = (T, R)) for
    the fused real-input product of layer 0, OR (in_image = the `img_out` buffer of the layer below (ring =
    planes_ring), w_ih [H,K], frames_rows) -- the spike-input product again, fetched with one bulk copy per frame.
    img_out (from `spike_image_buffer`, ring of img_ring frames, back-pressure bp_cnt / bp_target = the consumer's
    out_cnt / CTA count): also write this layer's spikes as the operand image of the layer above.
    Returns the bit-packed spike trace int32 [T,R,ceil(H/32)]."""
    lib, st = _prep(w_hh, bias, bn_scale, bn_shift, xproj, w_ih, out_h, out_c, out_hT, out_cT)
    H = w_hh.shape[1]
    if w_hh.shape[0] != H or bias.numel() != 2 * H:
        raise ValueError("recurrence_stream: shared gate weights only (w_hh [H,H], bias [2H])")
    if (xproj is not None) + (in_bits is not None) + (in_planes is not None) + (in_image is not None) != 1:
        raise ValueError("recurrence_stream: pass exactly one of xproj, (in_bits, w_ih), (in_planes, w_ih), "
                         "(in_image, w_ih)")
    if xproj is not None:
        T, R, _ = xproj.shape
        K_in = 0
    elif in_bits is not None:
        T, R, Wi = in_bits.shape
        K_in = w_ih.shape[1]
        if in_bits.dtype != torch.int32 or not in_bits.is_contiguous() or Wi != (K_in + 31) // 32 or w_ih.shape[0] != H:
            raise ValueError("recurrence_stream: in_bits / w_ih shapes")
    elif in_image is not None:
        T, R = frames_rows
        K_in = w_ih.shape[1]
        ring = T if (planes_ring <= 0 or planes_ring > T) else int(planes_ring)
        if w_ih.shape[0] != H or in_image.numel() * in_image.element_size() < lib.gsn_spike_image_bytes(ring, R, K_in):
            raise ValueError("recurrence_stream: in_image / w_ih shapes")
    else:
        T, R = frames_rows
        K_in = w_ih.shape[1]
        nt = lib.gsn_recurrence_stream_tile(R, H, K_in, 1, int(sm_budget))
        ring = T if (planes_ring <= 0 or planes_ring > T) else int(planes_ring)
        if w_ih.shape[0] != H or nt == 0 or in_planes.numel() * in_planes.element_size() < lib.gsn_xplanes_bytes(
                ring, R, K_in, nt):
            raise ValueError("recurrence_stream: in_planes / w_ih shapes")
    Wb = (H + 31) // 32
    if out_bits is None:
        out_bits = torch.empty((T, R, Wb), device=w_hh.device, dtype=torch.int32)
    elif out_bits.dtype != torch.int32 or not out_bits.is_contiguous() or tuple(out_bits.shape) != (T, R, Wb):
        raise ValueError("recurrence_stream: out_bits")
    for cnt in (in_cnt, out_cnt, bp_cnt):
        if cnt is not None and (cnt.dtype != torch.int32 or cnt.numel() != T or not cnt.is_contiguous()):
            raise ValueError("recurrence_stream: counters must be contiguous int32 [T]")
    if img_out is not None:
        ring_o = T if (img_ring <= 0 or img_ring > T) else int(img_ring)
        if img_out.numel() * img_out.element_size() < lib.gsn_spike_image_bytes(ring_o, R, H):
            raise ValueError("recurrence_stream: img_out too small")
    _lib.check(lib.gsn_recurrence_stream(
        _ptr(xproj), _ptr(in_bits), _ptr(in_planes), int(planes_ring), _ptr(w_ih), int(K_in), _ptr(w_hh), _ptr(bias),
        _ptr(bn_scale),
        _ptr(bn_shift), out_bits.data_ptr(), _ptr(out_h), _ptr(out_c), _ptr(out_hT), _ptr(out_cT), _ptr(in_cnt),
        int(in_target), _ptr(out_cnt), _ptr(spike_count), _ptr(in_image), _ptr(img_out), int(img_ring), _ptr(bp_cnt),
        int(bp_target), T, R, H, int(sm_budget), _ptr(workspace), st))
    LAUNCHES[0] += 1
    return out_bits


def spike_image_buffer(frames, R, H, device):
    """Operand-image buffer (ring of `frames` frames) a streaming recurrence fills through img_out for the layer above."""
    n = _lib.load().gsn_spike_image_bytes(int(frames), int(R), int(H))
    if n == 0:
        raise ValueError(f"spike_image_buffer: bad shape frames={frames} R={R} H={H}")
    return torch.zeros(n, device=device, dtype=torch.uint8)


def xplanes_buffer(T, R, K, nt, device):
    """Zeroed operand-image buffer of `xplanes_stream` for T frames (or a ring of T frames; uint8, 128-byte aligned by
    the caching allocator)."""
    n = _lib.load().gsn_xplanes_bytes(int(T), int(R), int(K), int(nt))
    if n == 0:
        raise ValueError(f"xplanes_buffer: bad shape T={T} R={R} K={K} nt={nt}")
    return torch.zeros(n, device=device, dtype=torch.uint8)


def xplanes_stream(cm, fb, N, lo, ctr, nbr, nt, xop, ln_weight=None, ln_bias=None, eps=1e-5, out_x=None, in_cnt=None,
                   in_target=0, out_cnt=None, ctas=1, ring=0, bp_cnt=None, bp_target=0, row_div=None):
    """Streaming gather + LayerNorm + bf16x3 split of a sequence model's layer-0 input into the B-operand images the
    fused layer-0 recurrence reads (gsn_xplanes_stream).  `xop` from `xplanes_buffer` (same nt)."""
    lib, st = _prep(cm, fb, ln_weight, ln_bias, out_x, row_div)
    T, B, f_cm = cm.shape
    K = ctr + 2 * nbr + (ctr if fb is not None else 0)
    div_mode = 0
    if row_div is not None:  # [B]: per utterance (offline laplace norm); [T, B*N]: per row and frame (cumulative)
        div_mode = 1 if row_div.numel() == B and row_div.dim() == 1 else 2
        if div_mode == 2 and tuple(row_div.shape) != (T, B * N):
            raise ValueError("xplanes_stream: row_div must be [B] or [T, B*N]")
    frames = T if (ring <= 0 or ring > T) else int(ring)
    if xop.dtype != torch.uint8 or xop.numel() < lib.gsn_xplanes_bytes(frames, B * N, K, nt) or xop.device != cm.device:
        raise ValueError("xplanes_stream: xop buffer")
    if out_x is not None and tuple(out_x.shape) != (T, B * N, K):
        raise ValueError("xplanes_stream: out_x shape")
    _lib.check(lib.gsn_xplanes_stream(_ptr(cm), f_cm, _ptr(fb), fb.shape[2] if fb is not None else 0, _ptr(ln_weight),
                                      _ptr(ln_bias), float(eps), _ptr(row_div), div_mode, _ptr(out_x), xop.data_ptr(),
                                      int(ring), _ptr(in_cnt),
                                      int(in_target), _ptr(out_cnt), _ptr(bp_cnt), int(bp_target), T, B, N, lo, ctr,
                                      nbr, int(nt), int(ctas), st))
    LAUNCHES[0] += 1
    return xop


def pre_stream(cm, fb, N, lo, ctr, nbr, w_ih, ln_weight=None, ln_bias=None, eps=1e-5, out_x=None, out_xproj=None,
               in_cnt=None, in_target=0, out_cnt=None, ctas_per_slice=1):
    """Streaming gather + LayerNorm + layer-0 input projection on tcgen05 (gsn_pre_stream).  Returns xproj [T,B*N,H]."""
    lib, st = _prep(cm, fb, w_ih, ln_weight, ln_bias, out_x, out_xproj)
    T, B, f_cm = cm.shape
    K = ctr + 2 * nbr + (ctr if fb is not None else 0)
    H = w_ih.shape[0]
    if w_ih.shape[1] != K:
        raise ValueError(f"pre_stream: w_ih{tuple(w_ih.shape)} vs K={K}")
    xproj = _out(out_xproj, (T, B * N, H), cm)
    if out_x is not None and tuple(out_x.shape) != (T, B * N, K):
        raise ValueError("pre_stream: out_x shape")
    _lib.check(lib.gsn_pre_stream(_ptr(cm), f_cm, _ptr(fb), fb.shape[2] if fb is not None else 0, _ptr(ln_weight),
                                  _ptr(ln_bias), float(eps), _ptr(w_ih), _ptr(out_x), _ptr(xproj), _ptr(in_cnt),
                                  int(in_target), _ptr(out_cnt), T, B, N, lo, ctr, nbr, H, int(ctas_per_slice), st))
    LAUNCHES[0] += 1
    return xproj


def linear_bits_stream(bits, w, bias=None, act=None, out=None, out_act=None, ctas=1, in_cnt=None, in_target=0,
                       out_cnt=None):
    """out[T,R,N] = spikes(bits [T,R,ceil(K/32)]) @ w[N,K]^T + bias as a streaming stage (gsn_linear_spike_bits_stream)."""
    lib, st = _prep(w, bias, out, out_act)
    T, R, Wk = bits.shape
    N, K = w.shape
    if bits.dtype != torch.int32 or not bits.is_contiguous() or Wk != (K + 31) // 32:
        raise ValueError("linear_bits_stream: bits shape")
    out = _out(out, (T, R, N), w)
    code = _ACT[act]
    out_act = _out(out_act, out.shape, w) if code else None
    _lib.check(lib.gsn_linear_spike_bits_stream(bits.data_ptr(), _ptr(w), _ptr(bias), _ptr(out), _ptr(out_act), code,
                                                T, R, K, N, int(ctas), _ptr(in_cnt), int(in_target), _ptr(out_cnt), st))
    LAUNCHES[0] += 1
    return (out, out_act) if code else out


def pick_backend(R, H, shared):
    return {1: "simt", 2: "tcgen05", 3: "tcgen05_i8"}[_lib.load().gsn_layer_recurrence_pick_backend(R, H, int(shared))]


def deepfilter_band(proj, spec_re, spec_im, out_re, out_im, N, ctr, df, S, lo):
    """Deep filter of one band straight from its proj output (MSF:315-346, layout MSF:160-167)."""
    lib, st = _prep(proj, spec_re, spec_im, out_re, out_im)
    T = proj.shape[0]
    B, F, _ = spec_re.shape
    F_out = out_re.shape[2]
    _lib.check(lib.gsn_deepfilter_band(_ptr(proj), _ptr(spec_re), _ptr(spec_im), _ptr(out_re), _ptr(out_im),
                                       T, B, N, ctr, df, S, lo, F, F_out, st))
    LAUNCHES[0] += 1


def _train_ws(R, H, shared, device):
    nbytes = _lib.load().gsn_layer_train_workspace_bytes(R, H, int(shared))
    ws = torch.empty((nbytes + 255) // 4 + 64, device=device, dtype=torch.float32)
    return ws[((-ws.data_ptr()) % 256) // 4:]


def layer_train_forward(xproj, w_hh, bias, bn_weight, bn_bias, running_mean, running_var, training, momentum, eps,
                        shared):
    """Training-path forward of one GSULayer (ESN:75-81 / 132-153, BatchNorm with batch statistics when
    `training`).  Returns (h, c, f, g, xhat, invstd); running statistics are updated in place."""
    lib, st = _prep(xproj, w_hh, bias, bn_weight, bn_bias, running_mean, running_var)
    T, R, gH = xproj.shape
    H = w_hh.shape[1]
    dev = xproj.device
    new = lambda *shape: torch.empty(shape, device=dev, dtype=torch.float32)  # noqa: E731
    h, c, f, g = new(T, R, H), new(T, R, H), new(T, R, H), new(T, R, H)
    batch_stats = bn_weight is not None and training
    xhat = new(T, R, H) if batch_stats else None
    invstd = new(T, H) if batch_stats else None
    if TC_TRAIN[0] and lib.gsn_layer_train_tc_supported(R, H, int(shared)):
        nbytes = lib.gsn_layer_train_tc_workspace_bytes(R, H)
        ws = torch.empty((nbytes + 255) // 4 + 64, device=dev, dtype=torch.float32)
        ws = ws[((-ws.data_ptr()) % 256) // 4:]
        rc = lib.gsn_layer_train_forward_tc(_ptr(xproj), _ptr(w_hh), _ptr(bias), _ptr(bn_weight), _ptr(bn_bias),
                                            _ptr(running_mean), _ptr(running_var), _ptr(h), _ptr(c), _ptr(f), _ptr(g),
                                            _ptr(xhat), _ptr(invstd), T, R, H, int(bool(training)), float(momentum),
                                            float(eps), 0, ws.data_ptr(), st)
        if rc != _lib.GSN_ENOSUP:
            _lib.check(rc)
            LAUNCHES[0] += 1
            return h, c, f, g, xhat, invstd
    ws = _train_ws(R, H, shared, dev)
    _lib.check(lib.gsn_layer_train_forward(_ptr(xproj), _ptr(w_hh), _ptr(bias), _ptr(bn_weight), _ptr(bn_bias),
                                           _ptr(running_mean), _ptr(running_var), _ptr(h), _ptr(c), _ptr(f), _ptr(g),
                                           _ptr(xhat), _ptr(invstd), T, R, H, int(shared), int(bool(training)),
                                           float(momentum), float(eps), ws.data_ptr(), st))
    LAUNCHES[0] += 2
    return h, c, f, g, xhat, invstd


def layer_train_backward(dh, w_hh, c, f, g, xhat, invstd, bn_weight, running_var, training, eps, shared):
    """BPTT of one GSULayer (surrogate gradient, BatchNorm backward).  Returns (dz [T,R,gH], dbias [2H],
    dgamma [H] | None, dbeta [H] | None)."""
    lib, st = _prep(dh, w_hh, c, f, g, xhat, invstd, bn_weight, running_var)
    T, R, H = dh.shape
    gH = w_hh.shape[0]
    dev = dh.device
    dz = torch.empty((T, R, gH), device=dev, dtype=torch.float32)
    nblocks = (R + 7) // 8
    dbias_part = torch.empty((nblocks, 2 * H), device=dev, dtype=torch.float32)
    batch_stats = bn_weight is not None and training
    dgamma = torch.empty(H, device=dev, dtype=torch.float32) if batch_stats else None
    dbeta = torch.empty(H, device=dev, dtype=torch.float32) if batch_stats else None
    ws = _train_ws(R, H, shared, dev)
    _lib.check(lib.gsn_layer_train_backward(_ptr(dh), _ptr(w_hh), _ptr(c), _ptr(f), _ptr(g), _ptr(xhat), _ptr(invstd),
                                            _ptr(bn_weight), _ptr(running_var), _ptr(dz), _ptr(dbias_part),
                                            _ptr(dgamma), _ptr(dbeta), T, R, H, int(shared), int(bool(training)),
                                            float(eps), ws.data_ptr(), st))
    LAUNCHES[0] += 1
    return dz, dbias_part.sum(dim=0), dgamma, dbeta
