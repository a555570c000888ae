from .logging import get_logger  # noqa: F401
from .metrics import JsonlSink, PrometheusSink, StepTimer, Throughput, measured_peaks, peak_tflops  # noqa: F401
