"""(reference: packages/prime-mcp-server/src/prime_mcp/tools/ssh.py:6-65)"""

from __future__ import annotations

from typing import Any

from ..client import call, make_prime_request


async def manage_ssh_keys(action: str = "list", key_name: str | None = None, public_key: str | None = None,
                          key_id: str | None = None, offset: int = 0, limit: int = 100) -> dict[str, Any]:  # fmt: skip
    if action == "list":
        return await call("GET", "ssh_keys/", "Unable to list SSH keys", params={"offset": max(0, offset), "limit": max(0, limit)})
    if action == "add":
        if not key_name or not public_key:
            return {"error": "key_name and public_key are required for adding SSH key"}
        return await call("POST", "ssh_keys/", "Unable to add SSH keys", json_data={"name": key_name, "publicKey": public_key})
    if action in ("delete", "set_primary"):
        if not key_id:
            return {"error": f"key_id is required for {'deleting' if action == 'delete' else 'setting primary'} SSH key"}
        if action == "delete":
            r = await make_prime_request("DELETE", f"ssh_keys/{key_id}")
            if r is not None and not r.get("error"):
                return {"success": True, "message": f"SSH key {key_id} deleted successfully"}
            return r or {"error": "Unable to delete SSH keys"}
        return await call("PATCH", f"ssh_keys/{key_id}", "Unable to set_primary SSH keys", json_data={"isPrimary": True})
    return {"error": f"Invalid action: {action}. Use 'list', 'add', 'delete', or 'set_primary'"}
