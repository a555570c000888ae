"""GPU / multi-node / disk availability (reference: packages/prime/src/prime_cli/api/availability.py:10-204).
Endpoints: GET /availability/{gpus,multi-node,disks,gpu-summary}, paginated with page/page_size."""

from __future__ import annotations

from typing import Any

from ._base import ApiModel, split_multi

PAGE_SIZE = 100


class CountSpec(ApiModel):
    min_count: int | None = None
    default_count: int | None = None
    max_count: int | None = None
    price_per_unit: float | None = None
    step: int | None = None
    default_included_in_price: bool | None = None
    additional_info: str | None = None


DiskConfig = ResourceConfig = CountSpec


class Prices(ApiModel):
    on_demand: float | None = None
    community_price: float | None = None
    is_variable: bool | None = None
    currency: str | None = None

    @property
    def price(self) -> float:
        for p in (self.community_price, self.on_demand):
            if p is not None:
                return p
        return float("inf")


class GPUAvailability(ApiModel):
    cloud_id: str
    gpu_type: str
    socket: str | None = None
    provider: str | None = None
    data_center: str | None = None
    country: str | None = None
    gpu_count: int
    gpu_memory: int
    disk: CountSpec
    vcpu: CountSpec
    memory: CountSpec
    internet_speed: float | None = None
    interconnect: int | None = None
    interconnect_type: str | None = None
    provisioning_time: int | None = None
    stock_status: str
    security: str | None = None
    prices: Prices
    images: list[str] | None = None
    is_spot: bool | None = None
    prepaid_time: int | None = None


class DiskAvailability(ApiModel):
    cloud_id: str | None = None
    provider: str | None = None
    data_center: str | None = None
    country: str | None = None
    region: str | None = None
    spec: CountSpec
    stock_status: str | None = None
    security: str | None = None
    is_multinode: bool | None = None


class AvailabilityClient:
    def __init__(self, client: Any, on_error=None) -> None:
        self.client = client
        self.on_error = on_error

    def _pages(self, url: str, params: dict[str, Any]) -> list[dict]:
        items: list[dict] = []
        page = 1
        while True:
            try:
                resp = self.client.get(url, params={**params, "page": page, "page_size": PAGE_SIZE})
            except Exception as e:  # partial results beat no results for a listing
                if self.on_error:
                    self.on_error(f"Error fetching availability page {page}: {e}")
                break
            batch = resp.get("items", [])
            items.extend(batch)
            if len(items) >= resp.get("totalCount", 0) or not batch:
                break
            page += 1
        return items

    def get(self, regions: list[str] | None = None, gpu_count: int | None = None, gpu_type: str | None = None,
            disks: list[str] | None = None) -> dict[str, list[GPUAvailability]]:  # fmt: skip
        params: dict[str, Any] = {}
        if split_multi(regions):
            params["regions"] = split_multi(regions)
        if gpu_count:
            params["gpu_count"] = str(gpu_count)
        if gpu_type:
            params["gpu_type"] = gpu_type
        if split_multi(disks):
            params["disks"] = split_multi(disks)
        grouped: dict[str, list[GPUAvailability]] = {}
        for url in ("/availability/gpus", "/availability/multi-node"):
            for raw in self._pages(url, params):
                grouped.setdefault(raw["gpuType"], []).append(GPUAvailability.model_validate(raw))
        return grouped

    def get_disks(self, regions: list[str] | None = None, data_center_id: str | None = None,
                  cloud_id: str | None = None) -> list[DiskAvailability]:  # fmt: skip
        params: dict[str, Any] = {}
        if regions:
            params["regions"] = split_multi(regions)
        if data_center_id:
            params["data_center_id"] = data_center_id
        if cloud_id:
            params["cloud_id"] = cloud_id
        return [DiskAvailability.model_validate(x) for x in self._pages("/availability/disks", params)]

    def get_available_gpu_types(self) -> list[str]:
        return list(self.client.get("/availability/gpu-summary").keys())
