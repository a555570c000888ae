from .availability import AvailabilityClient, DiskAvailability, GPUAvailability  # noqa: F401
from .deployments import Adapter, DeploymentsClient  # noqa: F401
from .disks import Disk, DisksClient  # noqa: F401
from .inference import InferenceAPIError, InferenceClient, InferencePaymentRequiredError  # noqa: F401
from .pods import Pod, PodsClient, PodStatus  # noqa: F401
from .rl import RLCheckpoint, RLClient, RLModel, RLRun  # noqa: F401
