"""Python handle over the C-ABI engine: one ``SnacEngine`` per GPU.

PyTorch is plumbing here (device tensors, streams); every number is produced by
``libsnacb.so``.  Replaces the module-global ``model`` + ``model.decode`` of
``/root/reference/Morpheus_Client/tts_engine/speechpipe.py:43-49,118``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib
from .weights import FoldedWeights

NoiseArg = Union[str, torch.Tensor, np.ndarray, None]
SAMPLES_PER_FRAME = 2048
SLICE_SAMPLES = 2048


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class SnacEngine:
    """Owns packed weights + workspace on one CUDA device; not re-entrant."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device: int = 0, precision: str = "fp16",
                 chunk_items: int = 0, trim: bool = True, fuse_ru: bool = True, lanes: int = 1,
                 persistent_ru: bool = False, fuse_convt_noise: bool = True, fuse_tail: bool = False,
                 persistent_convt: bool = True, fuse_ru256: bool = False, fuse_block: bool = True,
                 compose_convt_noise: bool = True):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        if not torch.cuda.is_available():
            raise _lib.SnacbError("no CUDA device: the SNAC decode path has no CPU fallback")
        self.device = int(device)
        self.torch_device = torch.device("cuda", self.device)
        self.precision = precision
        prec = {"fp32": _lib.PREC_FP32, "fp16": _lib.PREC_FP16, "fp16x3": _lib.PREC_FP16X3}[precision]
        cfg = _lib.Config(abi_version=_lib.ABI_VERSION, device=self.device, precision=prec,
                          chunk_items=int(chunk_items), trim=1 if trim else 0,
                          flags=(0 if fuse_ru else _lib.FLAG_NO_RU_FUSION) | (_lib.FLAG_PERSISTENT_RU if persistent_ru else 0)
                          | (0 if fuse_convt_noise else _lib.FLAG_NO_CONVT_NOISE_FUSION)
                          | (_lib.FLAG_TAIL_FUSION if fuse_tail else 0)
                          | (0 if persistent_convt else _lib.FLAG_NO_PERSISTENT_CONVT)
                          | (_lib.FLAG_FUSE_RU256 if fuse_ru256 else 0)
                          | (0 if fuse_block else _lib.FLAG_NO_BLOCK_FUSION)
                          | (0 if compose_convt_noise else _lib.FLAG_NO_CONVT_NOISE_COMPOSE), lanes=int(lanes))
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.zeros(1, device=self.torch_device)  # make sure the primary context exists
            rc = self._lib.snacb_create(C.byref(self._h), C.byref(cfg))
        if rc != _lib.OK:
            msg = self._lib.snacb_last_error(None)
            self._h = C.c_void_p()
            raise _lib.SnacbError(f"snacb_create failed ({rc}): {msg.decode() if msg else ''}")
        self._pin: Dict[str, torch.Tensor] = {}
        self.load_state_dict(state_dict)

    # ------------------------------------------------------------------ lifetime
    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            self._lib.snacb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str) -> None:
        _lib.check(self._lib, self._h, rc, what)

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        fw = sd if isinstance(sd, FoldedWeights) else FoldedWeights(sd)
        t = fw.tensors
        f32p = C.POINTER(C.c_float)

        def p(name):
            return C.cast(t[name].data_ptr(), f32p)

        w = _lib.Weights()
        for i in range(3):
            w.codebook[i], w.outproj_w[i], w.outproj_b[i] = p(f"codebook{i}"), p(f"outproj_w{i}"), p(f"outproj_b{i}")
        w.head_dw_w, w.head_dw_b, w.head_pw_w, w.head_pw_b = p("head_dw_w"), p("head_dw_b"), p("head_pw_w"), p("head_pw_b")
        for b in range(4):
            blk = w.block[b]
            blk.alpha, blk.convt_w, blk.convt_b, blk.noise_w = p(f"b{b}_alpha"), p(f"b{b}_convt_w"), p(f"b{b}_convt_b"), p(f"b{b}_noise_w")
            for r in range(3):
                ru = blk.ru[r]
                for fld in ("alpha1", "dw_w", "dw_b", "alpha2", "pw_w", "pw_b"):
                    setattr(ru, fld, p(f"b{b}_r{r}_{fld}"))
        w.tail_alpha, w.tail_w, w.tail_b = p("tail_alpha"), p("tail_w"), p("tail_b")
        self._check(self._lib.snacb_load_weights(self._h, C.byref(w)), "snacb_load_weights")
        self.has_encoder = bool(getattr(fw, "has_encoder", False))
        if self.has_encoder:  # encode direction
            ew = _lib.EncoderWeights()
            ew.in_w, ew.in_b, ew.out_dw_w, ew.out_dw_b = p("enc_in_w"), p("enc_in_b"), p("enc_out_dw_w"), p("enc_out_dw_b")
            for b in range(4):
                blk = ew.block[b]
                blk.alpha, blk.down_w, blk.down_b = p(f"enc{b}_alpha"), p(f"enc{b}_down_w"), p(f"enc{b}_down_b")
                for r in range(3):
                    for fld in ("alpha1", "dw_w", "dw_b", "alpha2", "pw_w", "pw_b"):
                        setattr(blk.ru[r], fld, p(f"enc{b}_r{r}_{fld}"))
            for i in range(3):
                ew.inproj_w[i], ew.inproj_b[i] = p(f"inproj_w{i}"), p(f"inproj_b{i}")
            self._check(self._lib.snacb_load_encoder_weights(self._h, C.byref(ew)), "snacb_load_encoder_weights")

    # ------------------------------------------------------------------ helpers
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.torch_device).cuda_stream

    @property
    def launch_count(self) -> int:
        return int(self._lib.snacb_launch_count(self._h))

    @property
    def graph_launch_count(self) -> int:
        return int(self._lib.snacb_graph_launch_count(self._h))

    @property
    def workspace_bytes(self) -> int:
        return int(self._lib.snacb_workspace_bytes(self._h))

    def _pinned(self, key: str, shape, dtype) -> torch.Tensor:
        n = int(np.prod(shape))
        buf = self._pin.get(key)
        if buf is None or buf.numel() < n or buf.dtype != dtype:
            buf = torch.empty(max(n, 1), dtype=dtype, pin_memory=True)
            self._pin[key] = buf
        return buf[:n].view(*shape)

    @staticmethod
    def _noise_mode(noise: NoiseArg) -> int:
        if noise is None or (isinstance(noise, str) and noise == "off"):
            return _lib.NOISE_OFF
        if isinstance(noise, str):
            if noise != "philox":
                raise ValueError(f"unknown noise mode {noise!r}")
            return _lib.NOISE_PHILOX
        return _lib.NOISE_TENSOR

    # ------------------------------------------------------------------ NS-1
    def deinterleave(self, tokens: torch.Tensor, ntok: Optional[Sequence[int]] = None, raw: bool = False,
                     max_frames: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        """tokens: cuda int32 [n, stride] -> (c0 [n,F], c1 [n,2F], c2 [n,4F], status [n]) on device."""
        assert tokens.is_cuda and tokens.dtype == torch.int32 and tokens.dim() == 2 and tokens.is_contiguous()
        n, stride = tokens.shape
        nt = None
        if ntok is not None:
            nt = np.ascontiguousarray(np.asarray(ntok, dtype=np.int32))
            assert nt.shape == (n,)
        mf = max_frames or max(1, (int(nt.max()) if nt is not None and n else stride) // 7)
        c0 = torch.empty((n, mf), dtype=torch.int32, device=tokens.device)
        c1 = torch.empty((n, 2 * mf), dtype=torch.int32, device=tokens.device)
        c2 = torch.empty((n, 4 * mf), dtype=torch.int32, device=tokens.device)
        status = torch.empty((n,), dtype=torch.int32, device=tokens.device)
        fn = self._lib.snacb_deinterleave_raw if raw else self._lib.snacb_deinterleave
        rc = fn(self._h, tokens.data_ptr(), stride, nt.ctypes.data if nt is not None else None, stride, n, mf,
                c0.data_ptr(), c1.data_ptr(), c2.data_ptr(), status.data_ptr(), self._stream())
        self._check(rc, "snacb_deinterleave")
        return c0, c1, c2, status

    # ------------------------------------------------------------------ streaming path
    def decode_windows_device(self, tokens: torch.Tensor, ntok: Optional[Sequence[int]] = None,
                              noise: NoiseArg = "philox", seed: int = 0, keys: Optional[Sequence[int]] = None,
                              pcm: Optional[torch.Tensor] = None, status: Optional[torch.Tensor] = None
                              ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Device-resident tick: tokens cuda int32 [n, stride] -> (pcm int16 [n,2048], status int32 [n])."""
        assert tokens.is_cuda and tokens.dtype == torch.int32 and tokens.dim() == 2 and tokens.is_contiguous()
        n, stride = tokens.shape
        if pcm is None:
            pcm = torch.empty((n, SLICE_SAMPLES), dtype=torch.int16, device=tokens.device)
        if status is None:
            status = torch.empty((n,), dtype=torch.int32, device=tokens.device)
        nt = None
        if ntok is not None:
            nt = np.ascontiguousarray(np.asarray(ntok, dtype=np.int32))
        mode = self._noise_mode(noise)
        nz_ptr, nz_stride = None, 0
        if mode == _lib.NOISE_TENSOR:
            nz = noise if isinstance(noise, torch.Tensor) else torch.from_numpy(np.asarray(noise, dtype=np.float32))
            nz = nz.to(device=tokens.device, dtype=torch.float32).contiguous()
            assert nz.dim() == 2 and nz.shape[0] == n
            self._keep = nz
            nz_ptr, nz_stride = nz.data_ptr(), nz.shape[1]
        kp = None
        if keys is not None:
            kp = np.ascontiguousarray(np.asarray(keys, dtype=np.uint64))
        rc = self._lib.snacb_decode_windows(self._h, tokens.data_ptr(), stride, nt.ctypes.data if nt is not None else None,
                                            stride, n, mode, nz_ptr, nz_stride, int(seed) & (2**64 - 1),
                                            kp.ctypes.data if kp is not None else None, pcm.data_ptr(),
                                            status.data_ptr(), self._stream())
        self._check(rc, "snacb_decode_windows")
        return pcm, status

    def decode_windows(self, tokens, ntok: Optional[Sequence[int]] = None, noise: NoiseArg = "philox", seed: int = 0,
                       keys: Optional[Sequence[int]] = None) -> Tuple[np.ndarray, np.ndarray]:
        """Host tick (the call ``convert_to_audio[_batch]`` makes): tokens int32 [n, stride] host array
        -> (pcm int16 [n,2048], status int32 [n]) numpy views of pinned buffers (valid until the next call)."""
        tok = np.ascontiguousarray(np.asarray(tokens, dtype=np.int32))
        assert tok.ndim == 2
        n, stride = tok.shape
        h_tok = self._pinned("tok", (n, stride), torch.int32)
        h_tok.numpy()[...] = tok
        h_pcm = self._pinned("pcm", (n, SLICE_SAMPLES), torch.int16)
        h_st = self._pinned("status", (n,), torch.int32)
        nt = None
        if ntok is not None:
            nt = np.ascontiguousarray(np.asarray(ntok, dtype=np.int32))
        mode = self._noise_mode(noise)
        nz_ptr, nz_stride = None, 0
        if mode == _lib.NOISE_TENSOR:
            nz = noise.detach().cpu().numpy() if isinstance(noise, torch.Tensor) else np.asarray(noise)
            nz = np.ascontiguousarray(nz, dtype=np.float32)
            assert nz.ndim == 2 and nz.shape[0] == n
            nz_ptr, nz_stride = nz.ctypes.data, nz.shape[1]
        kp = None
        if keys is not None:
            kp = np.ascontiguousarray(np.asarray(keys, dtype=np.uint64))
        with torch.cuda.device(self.device):
            rc = self._lib.snacb_decode_windows_host(self._h, h_tok.data_ptr(), stride,
                                                     nt.ctypes.data if nt is not None else None, stride, n, mode,
                                                     nz_ptr, nz_stride, int(seed) & (2**64 - 1),
                                                     kp.ctypes.data if kp is not None else None,
                                                     h_pcm.data_ptr(), h_st.data_ptr(), self._stream())
        self._check(rc, "snacb_decode_windows_host")
        return h_pcm.numpy(), h_st.numpy()

    def decode_windows_to_ring(self, ring, tokens, slots: Sequence[int], ntok: Optional[Sequence[int]] = None,
                               noise: NoiseArg = "philox", seed: int = 0, keys: Optional[Sequence[int]] = None,
                               eos: Optional[Sequence[int]] = None) -> Tuple[np.ndarray, np.ndarray]:
        """Host tick whose PCM goes straight into the pinned per-stream rings of ``ring`` (``egress.GpuPcmRing``): window i
        joins the ring of ``slots[i]`` (-1 = discard).  Returns (status int32 [n], emitted int32 [n]): the samples window i
        added to its ring (read them with ``ring.read(slot, nbytes)``), -1 where the ring had no room for a window (that
        window is dropped, the rest of the tick is unaffected)."""
        tok = np.ascontiguousarray(np.asarray(tokens, dtype=np.int32))
        assert tok.ndim == 2
        n, stride = tok.shape
        mode = self._noise_mode(noise)
        if mode == _lib.NOISE_TENSOR:
            raise ValueError("decode_windows_to_ring: injected noise is not supported (use decode_windows)")
        sl = np.ascontiguousarray(np.asarray(slots, dtype=np.int32))
        assert sl.shape == (n,)
        nt = np.ascontiguousarray(np.asarray(ntok, dtype=np.int32)) if ntok is not None else None
        kp = np.ascontiguousarray(np.asarray(keys, dtype=np.uint64)) if keys is not None else None
        eo = np.ascontiguousarray(np.asarray(eos, dtype=np.int32)) if eos is not None else None
        st = np.empty(n, dtype=np.int32)
        em = np.empty(n, dtype=np.int32)
        with torch.cuda.device(self.device):
            rc = self._lib.snacb_decode_windows_to_ring(
                self._h, ring.handle, tok.ctypes.data, stride, nt.ctypes.data if nt is not None else None, stride, n, mode,
                int(seed) & (2**64 - 1), kp.ctypes.data if kp is not None else None, sl.ctypes.data,
                eo.ctypes.data if eo is not None else None, st.ctypes.data, em.ctypes.data, self._stream())
        self._check(rc, "snacb_decode_windows_to_ring")
        return st, em

    def submit_windows(self, tokens, ntok: Optional[Sequence[int]] = None, noise: NoiseArg = "philox", seed: int = 0,
                       keys: Optional[Sequence[int]] = None) -> int:
        """Pipelined host tick: enqueue H2D + kernels + D2H and return a ticket at once.  ``wait_windows(ticket)``
        returns that tick's (pcm, status); submit tick t+1 before waiting for tick t and the copy-back and host work
        of t overlap the kernels of t+1 (two ticks in flight at most).  Noise "philox" or "off"."""
        tok = np.ascontiguousarray(np.asarray(tokens, dtype=np.int32))
        assert tok.ndim == 2
        n, stride = tok.shape
        mode = self._noise_mode(noise)
        if mode == _lib.NOISE_TENSOR:
            raise ValueError("submit_windows: injected noise is not supported (use decode_windows)")
        slot = getattr(self, "_pipe_slot", 0)
        h_pcm = self._pinned(f"pipe_pcm{slot}", (n, SLICE_SAMPLES), torch.int16)
        h_st = self._pinned(f"pipe_status{slot}", (n,), torch.int32)
        nt = np.ascontiguousarray(np.asarray(ntok, dtype=np.int32)) if ntok is not None else None
        kp = np.ascontiguousarray(np.asarray(keys, dtype=np.uint64)) if keys is not None else None
        ticket = C.c_int32(-1)
        with torch.cuda.device(self.device):
            rc = self._lib.snacb_decode_windows_host_submit(
                self._h, tok.ctypes.data, stride, nt.ctypes.data if nt is not None else None, stride, n, mode,
                int(seed) & (2**64 - 1), kp.ctypes.data if kp is not None else None, h_pcm.data_ptr(), h_st.data_ptr(),
                self._stream(), C.byref(ticket))
        self._check(rc, "snacb_decode_windows_host_submit")
        if not hasattr(self, "_pipe_out"):
            self._pipe_out = {}
        self._pipe_out[int(ticket.value)] = (h_pcm, h_st)
        self._pipe_slot = slot ^ 1
        return int(ticket.value)

    def wait_windows(self, ticket: int) -> Tuple[np.ndarray, np.ndarray]:
        """Block until the tick submitted under ``ticket`` is on the host: (pcm int16 [n,2048], status int32 [n]) as
        numpy views of pinned buffers (valid until the second next ``submit_windows``)."""
        with torch.cuda.device(self.device):
            rc = self._lib.snacb_decode_windows_host_wait(self._h, int(ticket))
        self._check(rc, "snacb_decode_windows_host_wait")
        h_pcm, h_st = self._pipe_out.pop(int(ticket))
        return h_pcm.numpy(), h_st.numpy()

    # ------------------------------------------------------------------ one-shot path
    def decode_codes(self, codes: Sequence[torch.Tensor], noise: NoiseArg = "philox", seed: int = 0,
                     want_pcm: bool = False):
        """``model.decode(codes)``: 3 tensors [B,F],[B,2F],[B,4F] -> float32 [B,1,2048F] (and int16 if asked)."""
        c = [x.to(device=self.torch_device, dtype=torch.int32).contiguous() for x in codes]
        B, F = c[0].shape
        if c[1].shape != (B, 2 * F) or c[2].shape != (B, 4 * F):
            raise ValueError("code tensors must be [B,F],[B,2F],[B,4F]")
        wav = torch.empty((B, 1, SAMPLES_PER_FRAME * F), dtype=torch.float32, device=self.torch_device)
        pcm = torch.empty((B, SAMPLES_PER_FRAME * F), dtype=torch.int16, device=self.torch_device) if want_pcm else None
        mode = self._noise_mode(noise)
        nz_ptr = None
        if mode == _lib.NOISE_TENSOR:
            nz = noise if isinstance(noise, torch.Tensor) else torch.from_numpy(np.asarray(noise, dtype=np.float32))
            nz = nz.to(device=self.torch_device, dtype=torch.float32).contiguous()
            assert tuple(nz.shape) == (B, _lib.NOISE_PER_FRAME * F)
            self._keep = nz
            nz_ptr = nz.data_ptr()
        with torch.cuda.device(self.device):
            rc = self._lib.snacb_decode_codes(self._h, c[0].data_ptr(), c[1].data_ptr(), c[2].data_ptr(), B, F, mode,
                                              nz_ptr, int(seed) & (2**64 - 1), wav.data_ptr(), _ptr(pcm), self._stream())
        self._check(rc, "snacb_decode_codes")
        self._keep_codes = c
        return (wav, pcm) if want_pcm else wav

    # ------------------------------------------------------------------ encode direction (N4)
    def encode(self, audio: torch.Tensor, return_latent: bool = False):
        """``model.encode(audio)`` of the published package: float audio ``[B,1,T]`` (or ``[B,T]``) -> the three code
        tensors ``[B,T'/4]``, ``[B,T'/2]``, ``[B,T']`` (int64, T' = padded length / 512); zero-padded on the right to a
        multiple of 2048 samples like ``SNAC.preprocess``."""
        if not getattr(self, "has_encoder", False):
            raise _lib.SnacbError("this engine was built from a state dict without encoder.* weights")
        x = audio.to(device=self.torch_device, dtype=torch.float32)
        if x.dim() == 3:
            x = x[:, 0, :]
        B, T = x.shape
        Tp = -(-T // 2048) * 2048
        if Tp != T:
            x = torch.nn.functional.pad(x, (0, Tp - T))
        x = x.contiguous()
        n = Tp // 512
        codes = [torch.empty((B, n // s), dtype=torch.int32, device=self.torch_device) for s in (4, 2, 1)]
        lat = torch.empty((B, n, 768), dtype=torch.float32, device=self.torch_device) if return_latent else None
        with torch.cuda.device(self.device):
            rc = self._lib.snacb_encode(self._h, x.data_ptr(), B, Tp, codes[0].data_ptr(), codes[1].data_ptr(), codes[2].data_ptr(),
                                        _ptr(lat), self._stream())
        self._check(rc, "snacb_encode")
        self._keep_audio = x
        out = [c.to(torch.int64) for c in codes]
        return (out, lat.transpose(1, 2)) if return_latent else out

    def fill_noise(self, seed: int, n_win: int, frames: int, keys: Optional[Sequence[int]] = None) -> torch.Tensor:
        out = torch.empty((n_win, _lib.NOISE_PER_FRAME * frames), dtype=torch.float32, device=self.torch_device)
        kp = np.ascontiguousarray(np.asarray(keys, dtype=np.uint64)) if keys is not None else None
        with torch.cuda.device(self.device):
            rc = self._lib.snacb_fill_noise(self._h, int(seed) & (2**64 - 1), kp.ctypes.data if kp is not None else None,
                                            n_win, frames, out.data_ptr(), out.shape[1], self._stream())
        self._check(rc, "snacb_fill_noise")
        return out

    # ------------------------------------------------------------------ measurement
    def profile(self, on: bool) -> None:
        self._check(self._lib.snacb_profile_enable(self._h, 1 if on else 0), "snacb_profile_enable")

    def profile_read(self) -> Dict[str, Dict[str, float]]:
        """Per-kernel-class totals since ``profile(True)``: launches, device ms, executed flops, bytes."""
        rows = (_lib.KernelStat * 16)()
        n = self._lib.snacb_profile_read(self._h, rows, 16)
        if n < 0:
            self._check(n, "snacb_profile_read")
        return {rows[i].name.decode(): {"launches": int(rows[i].launches), "ms": float(rows[i].ms),
                                        "flops": float(rows[i].flops), "bytes": float(rows[i].bytes)}
                for i in range(min(n, 16))}

    # ------------------------------------------------------------------ bring-up taps
    def set_tap(self, stage: int, capacity_floats: int = 0) -> Optional[torch.Tensor]:
        if stage < 0:
            self._lib.snacb_set_tap(self._h, -1, None, 0)
            self._tap = None
            return None
        self._tap = torch.zeros(capacity_floats, dtype=torch.float32, device=self.torch_device)
        self._check(self._lib.snacb_set_tap(self._h, stage, self._tap.data_ptr(), capacity_floats), "snacb_set_tap")
        return self._tap

    def get_tap(self) -> Tuple[torch.Tensor, int]:
        r, ch, lo, it = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        self._lib.snacb_get_tap_shape(self._h, C.byref(r), C.byref(ch), C.byref(lo), C.byref(it))
        n = r.value * ch.value * it.value
        torch.cuda.synchronize(self.torch_device)
        return self._tap[:n].view(it.value, r.value, ch.value).clone(), lo.value
