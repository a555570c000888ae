import pytest


@pytest.fixture(autouse=True)
def isolated_home(tmp_path, monkeypatch):
    """Every test gets its own ~/.prime (the reference isolates per xdist worker: prime-sandboxes/tests/conftest.py:12-28)."""
    home = tmp_path / "home"
    home.mkdir()
    monkeypatch.setenv("HOME", str(home))
    monkeypatch.setattr("pathlib.Path.home", lambda: home)
    for k in ("PRIME_API_KEY", "PRIME_TEAM_ID", "PRIME_USER_ID", "PRIME_API_BASE_URL", "PRIME_BASE_URL", "PRIME_CONTEXT",
              "PRIME_FRONTEND_URL", "PRIME_INFERENCE_URL", "PRIME_SSH_KEY_PATH"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv("PRIME_DISABLE_VERSION_CHECK", "1")
    monkeypatch.setenv("COLUMNS", "200")
    return home


@pytest.fixture
def anyio_backend():
    return "asyncio"
