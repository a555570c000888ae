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


class NvlinkCounters:
    """Cumulative NVLink payload counters of one GPU, read through NVML field values
    (``NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX`` / ``_RX``, unit KiB; scope ``UINT_MAX`` = all links, else per link).

    ``read()`` → {"tx_bytes", "rx_bytes", "per_link": [[tx, rx], …]} or ``None`` when NVML, the field or the GPU's NVLink is not
    available. Two reads around a timed region give the bytes that crossed NVLink in it — the raw counters, nothing derived; they
    cover everything on the links (this job's peer loads / stores, NCCL of the comparison arm, other tenants of the box)."""

    ALL_LINKS = 0xFFFFFFFF
    MAX_LINKS = 18

    def __init__(self, index: int, nvml=None):
        self.ok, self.h, self.nv = False, None, nvml
        try:
            if self.nv is None:
                import pynvml

                pynvml.nvmlInit()
                self.nv = pynvml
                try:
                    import torch

                    uuid = "GPU-" + str(torch.cuda.get_device_properties(torch.cuda.current_device()).uuid)
                    self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
                except Exception:  # noqa: BLE001
                    self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            else:
                self.h = self.nv.nvmlDeviceGetHandleByIndex(index)
            self.tx, self.rx = self.nv.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, self.nv.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX
            self.ok = True
        except Exception:  # noqa: BLE001
            self.ok = False

    @staticmethod
    def _value(fv) -> int | None:
        """One ``nvmlFieldValue_t`` → integer, or None when NVML reports an error for that field."""
        if getattr(fv, "nvmlReturn", 1) != 0:
            return None
        v, t = fv.value, getattr(fv, "valueType", 3)
        # NVML_VALUE_TYPE_*: 0 double · 1 unsigned int · 2 unsigned long · 3 unsigned long long · 4 signed long long
        return int({0: getattr(v, "dVal", 0), 1: getattr(v, "uiVal", 0), 2: getattr(v, "ulVal", 0), 3: getattr(v, "ullVal", 0),
                    4: getattr(v, "sllVal", 0)}.get(t, getattr(v, "ullVal", 0)))  # fmt: skip

    def read(self) -> dict | None:
        if not self.ok:
            return None
        try:
            tot = self.nv.nvmlDeviceGetFieldValues(self.h, [(self.tx, self.ALL_LINKS), (self.rx, self.ALL_LINKS)])
            tx, rx = self._value(tot[0]), self._value(tot[1])
            per_link = []
            for link in range(self.MAX_LINKS):
                vals = self.nv.nvmlDeviceGetFieldValues(self.h, [(self.tx, link), (self.rx, link)])
                a, b = self._value(vals[0]), self._value(vals[1])
                if a is None and b is None:
                    break
                per_link.append([(a or 0) * 1024, (b or 0) * 1024])
            if tx is None or rx is None:  # no aggregate scope on this driver: sum the links
                if not per_link:
                    return None
                tx, rx = sum(x[0] for x in per_link) // 1024, sum(x[1] for x in per_link) // 1024
            return {"tx_bytes": tx * 1024, "rx_bytes": rx * 1024, "per_link": per_link}
        except Exception:  # noqa: BLE001
            return None

    @staticmethod
    def delta(before: dict | None, after: dict | None, steps: int) -> dict | None:
        """Bytes per step between two reads (None when either read failed)."""
        if not before or not after or steps <= 0:
            return None
        links = [[(a[0] - b[0]) // steps, (a[1] - b[1]) // steps] for a, b in zip(after["per_link"], before["per_link"])]
        return {"tx_bytes_per_step": (after["tx_bytes"] - before["tx_bytes"]) // steps, "rx_bytes_per_step": (after["rx_bytes"] - before["rx_bytes"]) // steps,
                "links_active": sum(1 for x in links if x[0] or x[1]), "per_link_bytes_per_step": links,
                "source": "NVML NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX/RX (KiB counters, all links), rank 0's GPU, raw"}  # fmt: skip
