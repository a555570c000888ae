"""Pure helpers behind ``prime env``: what goes into a source archive, the deterministic content hash, version bumping,
wheel metadata, safe tar extraction, slug parsing and install-command construction
(reference: packages/prime/src/prime_cli/commands/env.py:113-160, :452-566, :1717-1910, :2652-2735, :2882-2968)."""

from __future__ import annotations

import hashlib
import os
import re
import shutil
import subprocess
import sys
import tarfile
import zipfile
from pathlib import Path
from typing import Callable, Iterable

MAX_TARBALL_SIZE_LIMIT = 250 * 1024 * 1024  # soft limit: warn above 250 MB
SKIP_DIR_NAMES = {"dist", "__pycache__", "build", "outputs"}
TOP_LEVEL_PATTERNS = ("README.md", "pyproject.toml", "*.py")


# --------------------------------------------------------------------------------------------------- archive contents
def include_file(path: Path, base: Path) -> bool:
    if not path.is_file() or path.is_symlink() or path.name.startswith("."):
        return False  # symlinks would be refused by safe_tar_extract on the other side
    return "__pycache__" not in path.relative_to(base).parts


def include_directory(path: Path) -> bool:
    return path.is_dir() and not path.name.startswith(".") and path.name not in SKIP_DIR_NAMES and not path.name.endswith(".egg-info")


def gitignore_matcher(env_path: Path) -> Callable[[str], bool] | None:
    gi = env_path / ".gitignore"
    if not gi.exists():
        return None
    from gitignore_parser import parse_gitignore

    return parse_gitignore(str(gi), base_dir=str(env_path))


def collect_archive_files(env_path: Path) -> list[Path]:
    """Deterministic (sorted by posix relative path) list of files that define an environment: top-level
    README/pyproject/*.py plus every non-ignored file below non-ignored sub-directories, honouring the root .gitignore."""
    ignored = gitignore_matcher(env_path) or (lambda _p: False)
    found: dict[str, Path] = {}

    def add(p: Path) -> None:
        if include_file(p, env_path) and not ignored(str(p)):
            found[p.relative_to(env_path).as_posix()] = p

    for pattern in TOP_LEVEL_PATTERNS:
        for p in sorted(env_path.glob(pattern), key=lambda x: x.name):
            add(p)
    prune = lambda d: not include_directory(d) or ignored(str(d))  # noqa: E731
    for sub in sorted(env_path.iterdir(), key=lambda x: x.name):
        if prune(sub):
            continue
        for root, dirs, files in os.walk(sub):
            rp = Path(root)
            dirs[:] = sorted(d for d in dirs if not prune(rp / d))
            for f in sorted(files):
                add(rp / f)
    return [found[k] for k in sorted(found)]


def compute_content_hash(env_path: Path) -> str:
    h = hashlib.sha256()
    for p in collect_archive_files(env_path):
        h.update(f"file:{p.relative_to(env_path).as_posix()}".encode())
        try:
            h.update(p.read_bytes())
        except OSError:
            pass
    return h.hexdigest()


def build_source_archive(env_path: Path, dest: Path) -> int:
    with tarfile.open(dest, "w:gz") as tar:
        for p in collect_archive_files(env_path):
            tar.add(p, arcname=str(p.relative_to(env_path)))
    return dest.stat().st_size


def has_environment_code(env_path: Path) -> bool:
    return any(env_path.glob("*.py")) or any((d / "__init__.py").exists() for d in env_path.iterdir() if d.is_dir())


# --------------------------------------------------------------------------------------------------- slugs / versions
def parse_environment_slug(environment: str) -> tuple[str, str]:
    owner, sep, name = environment.partition("/")
    if not sep or not owner or not name or "/" in name:
        raise ValueError(f"Invalid environment '{environment}': expected owner/name")
    return owner, name.split("@", 1)[0]


def validate_env_id(env_id: str) -> tuple[str, str]:
    """``owner/name[@version]`` → ("owner/name", version or "latest")."""
    base, sep, version = env_id.partition("@")
    if sep and not version:
        raise ValueError("version cannot be empty after '@'")
    parts = base.split("/")
    if len(parts) != 2 or not all(parts):
        raise ValueError("expected format owner/name or owner/name@version")
    return base, version or "latest"


def normalize_package_name(name: str) -> str:
    return re.sub(r"[-_.]+", "-", name).lower()


_VER = re.compile(r"^(\d+)\.(\d+)\.(\d+)(.*)$")


def bump_version(version: str) -> str:
    """1.2.3 → 1.2.4 (any pre/post suffix is dropped)."""
    m = _VER.match(version)
    if not m:
        raise ValueError(f"cannot bump non-semver version {version!r}")
    return f"{m.group(1)}.{m.group(2)}.{int(m.group(3)) + 1}"


def bump_rc_version(version: str) -> str:
    """1.2.3 → 1.2.3rc1 → 1.2.3rc2."""
    m = re.match(r"^(.*?)rc(\d+)$", version)
    return f"{m.group(1)}rc{int(m.group(2)) + 1}" if m else f"{version}rc1"


def bump_post_version(version: str) -> str:
    """1.2.3 → 1.2.3.post1 → 1.2.3.post2."""
    m = re.match(r"^(.*?)\.post(\d+)$", version)
    return f"{m.group(1)}.post{int(m.group(2)) + 1}" if m else f"{version}.post1"


def update_pyproject_version(pyproject: Path, new_version: str) -> None:
    """Rewrite only the ``version = "..."`` line inside [project] (keeps comments and formatting)."""
    lines = pyproject.read_text().splitlines(keepends=True)
    in_project = False
    for i, line in enumerate(lines):
        s = line.strip()
        if s.startswith("["):
            in_project = s == "[project]"
        elif in_project and re.match(r"version\s*=", s):
            lines[i] = re.sub(r'(version\s*=\s*)["\'][^"\']*["\']', rf'\g<1>"{new_version}"', line)
            pyproject.write_text("".join(lines))
            return
    raise ValueError("no version field found in [project]")


def extract_requires_dist_from_wheel(wheel: Path) -> list[str]:
    try:
        with zipfile.ZipFile(wheel) as z:
            meta = next((n for n in z.namelist() if n.endswith(".dist-info/METADATA")), None)
            if not meta:
                return []
            return [ln.split(":", 1)[1].strip() for ln in z.read(meta).decode("utf-8", "replace").splitlines() if ln.startswith("Requires-Dist:")]
    except (zipfile.BadZipFile, OSError):
        return []


# --------------------------------------------------------------------------------------------------- tar safety
def safe_tar_extract(tar: tarfile.TarFile, dest: Path) -> None:
    """Refuse symlinks, hardlinks, absolute paths and anything resolving outside ``dest``; only then extract."""
    dest = dest.resolve()
    for m in tar.getmembers():
        p = Path(m.name)
        if m.issym() or m.islnk():
            raise ValueError(f"Refusing to extract {'symlink' if m.issym() else 'hardlink'}: {m.name}")
        if p.is_absolute():
            raise ValueError(f"Refusing to extract absolute path: {m.name}")
        if ".." in p.parts:
            raise ValueError(f"Refusing to extract path with '..': {m.name}")
        if not (dest / p).resolve().is_relative_to(dest):
            raise ValueError(f"Path escapes destination directory: {m.name}")
    tar.extractall(dest, filter="data")


def validate_path_component(component: str, what: str) -> None:
    """``owner`` / ``name`` / ``version`` become directory names under the wheel cache: each must be ONE harmless path segment."""
    problems = (
        (not component, "cannot be empty"),
        ("\x00" in component, "cannot contain null bytes"),
        (".." in component, "cannot contain '..'"),
        ("/" in component or "\\" in component, "cannot contain path separators"),
    )
    for bad, why in problems:
        if bad:
            raise ValueError(f"Invalid {what}: {why}")


def env_cache_dir() -> Path:
    d = Path.home() / ".prime" / "wheel_cache"
    d.mkdir(parents=True, exist_ok=True)
    return d


# --------------------------------------------------------------------------------------------------- install commands
def uv_pip_command(sub: str, *args: str) -> list[str]:
    """``uv pip`` targeting THIS interpreter's environment."""
    return ["uv", "pip", sub, "--python", sys.executable, *args]


def is_valid_url(url: str) -> bool:
    return bool(re.match(r"https?://[^\s/]+", url or ""))


def process_wheel_url(url: str | None) -> str | None:
    return url if url and is_valid_url(url) else None


def get_install_command(tool: str, wheel_url: str, package: str, no_upgrade: bool = False) -> list[str]:
    if not is_valid_url(wheel_url):
        raise ValueError(f"Invalid wheel URL: {wheel_url}")
    if tool == "uv":
        return uv_pip_command("install", *([] if no_upgrade else ["-P", package]), wheel_url)
    if tool == "pip":
        return ["pip", "install", *([] if no_upgrade else ["--upgrade"]), wheel_url]
    raise ValueError(f"Unsupported package manager: {tool}")


def build_install_command(name: str, version: str, simple_index_url: str | None, wheel_url: str | None, tool: str = "uv",
                          no_upgrade: bool = False, url_dependencies: Iterable[str] | None = None) -> list[str] | None:  # fmt: skip
    """Prefer the hub's simple index (lets the resolver see every version), else the direct wheel URL."""
    pkg = normalize_package_name(name)
    deps = list(url_dependencies or [])
    if simple_index_url:
        spec = f"{pkg}=={version}" if version and version != "latest" else pkg
        if tool == "uv":
            return uv_pip_command("install", *([] if no_upgrade else ["-P", pkg]), spec, *deps, "--extra-index-url", simple_index_url)
        return ["pip", "install", *([] if no_upgrade else ["--upgrade"]), spec, *deps, "--extra-index-url", simple_index_url]
    if wheel_url:
        try:
            return get_install_command(tool, wheel_url, pkg, no_upgrade) + deps
        except ValueError:
            return None
    return None


def is_environment_installed(env_name: str, required_version: str | None = None, runner=subprocess.run) -> bool:
    try:
        r = runner(uv_pip_command("show", normalize_package_name(env_name)), capture_output=True, text=True)
    except Exception:
        return False
    if r.returncode != 0:
        return False
    if required_version and required_version != "latest":
        for line in r.stdout.splitlines():
            if line.startswith("Version:"):
                return line.split(":", 1)[1].strip() == required_version
        return False
    return True


def build_wheel(env_path: Path) -> Path:
    """Fresh wheel under ``dist/`` via ``uv build`` (preferred) or ``python -m build``."""
    dist = env_path / "dist"
    if dist.exists():
        shutil.rmtree(dist)
    if shutil.which("uv"):
        subprocess.run(["uv", "build", "--wheel", "--out-dir", "dist"], cwd=env_path, capture_output=True, text=True, check=True)
    else:
        subprocess.run([sys.executable, "-m", "build", "--wheel", str(env_path)], capture_output=True, text=True, check=True)
    wheels = list(dist.glob("*.whl"))
    if not wheels:
        raise FileNotFoundError("No wheel file found after build")
    return wheels[0]
