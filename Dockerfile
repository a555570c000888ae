# CLI image: the platform layer only (no CUDA toolchain) — `docker run … prime-b200 pods list`.
# The training engine is built on the GPU host with `python -c "import __graft_entry__ as g; g.build()"`.
ARG PYTHON_VERSION=3.12
FROM python:${PYTHON_VERSION}-slim AS build
COPY --from=ghcr.io/astral-sh/uv:latest /uv /usr/local/bin/uv
WORKDIR /src
COPY pyproject.toml README.md ./
COPY prime_b200 ./prime_b200
COPY diloco ./diloco
RUN uv build --wheel --out-dir /dist

FROM python:${PYTHON_VERSION}-slim
COPY --from=ghcr.io/astral-sh/uv:latest /uv /usr/local/bin/uv
COPY --from=build /dist/*.whl /tmp/
# CPU torch keeps the image small; the CLI never touches a GPU
RUN uv pip install --system --index-url https://download.pytorch.org/whl/cpu torch \
 && uv pip install --system "$(ls /tmp/*.whl)[mcp,rpc]" && rm /tmp/*.whl \
 && useradd --create-home prime
USER prime
ENV PRIME_DISABLE_VERSION_CHECK=1
ENTRYPOINT ["prime-b200"]
CMD ["--help"]
