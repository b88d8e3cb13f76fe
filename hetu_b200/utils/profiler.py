"""Device-level profiler: memory (NVML device view + allocator reserved / allocated / peak), NVLink TX/RX counters between
two points, per-micro-batch memory log, SM clocks / throttle reasons.
(ref: hetu/graph/profiler.{h,cc} CUDAProfiler + CUDAMemoryInfo, HETU_MEMORY_PROFILE / HETU_MEMORY_LOG_FILE handling in
executable_graph.cc)"""
from __future__ import annotations

import json
import os
import time
from dataclasses import asdict, dataclass, field
from typing import Dict, List, Optional

import torch

try:
    import pynvml
    _HAVE_NVML = True
except Exception:   # noqa: BLE001
    _HAVE_NVML = False


@dataclass
class CUDAMemoryInfo:
    mempool_reserved: int = 0
    mempool_peak_reserved: int = 0
    mempool_allocated: int = 0
    all_reserved: int = 0          # device-wide used bytes (NVML): includes other processes, NCCL, contexts
    limit: int = 0


@dataclass
class MicroBatchMemoryInfo:
    is_forward: bool
    stage_id: int
    micro_batch_id: int
    begin: CUDAMemoryInfo = field(default_factory=CUDAMemoryInfo)
    end: CUDAMemoryInfo = field(default_factory=CUDAMemoryInfo)


class CUDAProfiler:
    def __init__(self, device: Optional[int] = None, log_file: Optional[str] = None):
        self.device = torch.cuda.current_device() if (device is None and torch.cuda.is_available()) else (device or 0)
        self.log_file = log_file or os.environ.get("HETU_MEMORY_LOG_FILE")
        self._nvml = None
        if _HAVE_NVML and torch.cuda.is_available():
            try:
                pynvml.nvmlInit()
                idx = int(os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[self.device]) if os.environ.get("CUDA_VISIBLE_DEVICES") else self.device
                self._nvml = pynvml.nvmlDeviceGetHandleByIndex(idx)
            except Exception:   # noqa: BLE001
                self._nvml = None
        self._nvlink_start: Optional[Dict[str, int]] = None
        self.micro_batch_log: List[MicroBatchMemoryInfo] = []

    # ---- memory
    def get_current_memory_info(self) -> CUDAMemoryInfo:
        info = CUDAMemoryInfo()
        if torch.cuda.is_available():
            info.mempool_reserved = torch.cuda.memory_reserved(self.device)
            info.mempool_peak_reserved = torch.cuda.max_memory_reserved(self.device)
            info.mempool_allocated = torch.cuda.memory_allocated(self.device)
            free, total = torch.cuda.mem_get_info(self.device)
            info.all_reserved, info.limit = total - free, total
        if self._nvml is not None:
            try:
                m = pynvml.nvmlDeviceGetMemoryInfo(self._nvml)
                info.all_reserved, info.limit = int(m.used), int(m.total)
            except Exception:   # noqa: BLE001
                pass
        return info

    def print_current_memory_info(self, prefix: str = ""):
        i = self.get_current_memory_info()
        mb = 1 << 20
        print(f"{prefix} device {self.device}: pool reserved {i.mempool_reserved // mb} MiB (peak {i.mempool_peak_reserved // mb}), "
              f"allocated {i.mempool_allocated // mb} MiB, device used {i.all_reserved // mb} / {i.limit // mb} MiB", flush=True)

    def record_micro_batch(self, is_forward: bool, stage_id: int, micro_batch_id: int, begin: CUDAMemoryInfo, end: CUDAMemoryInfo):
        rec = MicroBatchMemoryInfo(is_forward, stage_id, micro_batch_id, begin, end)
        self.micro_batch_log.append(rec)
        if self.log_file:
            with open(self.log_file, "a") as f:
                f.write(json.dumps(asdict(rec)) + "\n")

    # ---- NVLink
    def _nvlink_counters(self) -> Dict[str, int]:
        tx = rx = 0
        links = 0
        if self._nvml is not None:
            for link in range(18):
                try:
                    if pynvml.nvmlDeviceGetNvLinkState(self._nvml, link) != pynvml.NVML_FEATURE_ENABLED:
                        continue
                    links += 1
                    fv = pynvml.nvmlDeviceGetFieldValues(self._nvml, [(pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, link),
                                                                      (pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX, link)])
                    tx += int(fv[0].value.ullVal)
                    rx += int(fv[1].value.ullVal)
                except Exception:   # noqa: BLE001
                    continue
        return {"tx_kib": tx, "rx_kib": rx, "links": links, "t": time.perf_counter()}

    def profile_nvlink_start(self):
        self._nvlink_start = self._nvlink_counters()

    def profile_nvlink_end(self) -> Dict[str, float]:
        end, start = self._nvlink_counters(), self._nvlink_start or self._nvlink_counters()
        dt = max(end["t"] - start["t"], 1e-9)
        tx, rx = (end["tx_kib"] - start["tx_kib"]) * 1024, (end["rx_kib"] - start["rx_kib"]) * 1024
        return {"links": end["links"], "tx_bytes": tx, "rx_bytes": rx, "tx_gbs": tx / dt / 1e9, "rx_gbs": rx / dt / 1e9, "seconds": dt}

    # ---- clocks
    def clocks(self) -> Dict[str, object]:
        out: Dict[str, object] = {}
        if self._nvml is not None:
            try:
                out["sm_mhz"] = pynvml.nvmlDeviceGetClockInfo(self._nvml, pynvml.NVML_CLOCK_SM)
                out["sm_max_mhz"] = pynvml.nvmlDeviceGetMaxClockInfo(self._nvml, pynvml.NVML_CLOCK_SM)
                r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(self._nvml) if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(self._nvml)
                names = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
                out["reasons"] = [k for k, bit in names.items() if r & bit]
                out["power_w"] = pynvml.nvmlDeviceGetPowerUsage(self._nvml) / 1000.0
            except Exception:   # noqa: BLE001
                pass
        return out


_profilers: Dict[int, CUDAProfiler] = {}


def get_cuda_profiler(device: Optional[int] = None) -> CUDAProfiler:
    key = -1 if device is None else device
    if key not in _profilers:
        _profilers[key] = CUDAProfiler(device)
    return _profilers[key]
