"""GPU clock / throttle sampling through NVML while a timed region runs (the recipe's "clocks line").

Used by ``bench.py`` (JSON ``clocks`` key) and by the trainer's monitor (``monitor.clocks = true``).
"""

from __future__ import annotations

import threading

class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of this process's GPU through NVML during the timed region."""

    def __init__(self, index: int, period_s: float = 0.2):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples: list[int] = []
        self.reasons: set[str] = set()
        self.max_mhz = 0
        self.power: list[float] = []
        self._stop = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                import torch

                uuid = "GPU-" + str(torch.cuda.get_device_properties(torch.cuda.current_device()).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self) -> None:
        if not self.ok:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(self.period)

    def finish(self) -> dict:
        self._stop.set()
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz or None, "reasons": ["nvml_unavailable"]}
        s = sorted(self.samples)
        return {
            "sm_mhz": s[len(s) // 2],
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "power_w_max": round(max(self.power), 1) if self.power else None,
            "samples": len(s),
        }
