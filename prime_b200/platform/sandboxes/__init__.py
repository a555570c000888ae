"""Remote code-execution sandbox SDK (sync + async)."""

from ..core.client import APIError, APITimeoutError, PaymentRequiredError, UnauthorizedError  # noqa: F401
from .client import APIClient, AsyncAPIClient  # noqa: F401  (the shared clients with the SDK's retry default)
from ..core.config import Config  # noqa: F401
from .exceptions import (  # noqa: F401
    CommandTimeoutError,
    DownloadTimeoutError,
    SandboxFileNotFoundError,
    SandboxImagePullError,
    SandboxNotRunningError,
    SandboxOOMError,
    SandboxTimeoutError,
    UploadTimeoutError,
)
from .models import (  # noqa: F401
    AdvancedConfigs,
    BackgroundJob,
    BackgroundJobStatus,
    BulkDeleteSandboxRequest,
    BulkDeleteSandboxResponse,
    CommandRequest,
    CommandResponse,
    CreateSandboxRequest,
    DockerImageCheckResponse,
    ExposedPort,
    ExposePortRequest,
    FileUploadResponse,
    ListExposedPortsResponse,
    ReadFileResponse,
    RegistryCredentialSummary,
    Sandbox,
    SandboxListResponse,
    SandboxStatus,
    SSHSession,
    UpdateSandboxRequest,
)
from .sandbox import AsyncSandboxClient, AsyncTemplateClient, SandboxClient, TemplateClient  # noqa: F401

__version__ = "0.1.0"

TimeoutError = APITimeoutError  # the reference exports the transport timeout under this short name as well
