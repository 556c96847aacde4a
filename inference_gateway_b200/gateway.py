"""ctypes view of the host-side mirror (include/sse_gateway.h): a provider whose StreamChatCompletions returns a
channel of lines, the batcher tick, and the MCP agent's per-iteration view. Plumbing only."""
from __future__ import annotations

import ctypes as C

from . import _abi as A


class Gateway:
    def __init__(self, device: int = 0, max_conns: int = 256, bytes_per_batch: int = 1 << 20):
        self.L = L = A.load()
        vp, i32 = C.c_void_p, C.c_int
        L.ssegw_new.argtypes = [i32, C.c_uint32, C.c_uint32, C.POINTER(i32)]
        L.ssegw_new.restype = vp
        L.ssegw_free.argtypes = [vp]
        L.ssegw_free.restype = None
        L.ssegw_stream_chat_completions.argtypes = [vp, C.c_uint8]
        L.ssegw_upstream_write.argtypes = [vp, i32, C.c_char_p, C.c_size_t]
        L.ssegw_upstream_write.restype = C.c_size_t
        L.ssegw_upstream_close.argtypes = [vp, i32]
        L.ssegw_upstream_close.restype = None
        L.ssegw_pump.argtypes = [vp]
        L.ssegw_recv.argtypes = [vp, i32, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.ssegw_agent_recv.argtypes = [vp, i32, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.ssegw_proxy_stream.argtypes = [vp]
        L.ssegw_proxy_step.argtypes = [vp, i32, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.ssegw_release_stream.argtypes = [vp, i32]
        L.ssegw_release_stream.restype = None
        L.ssegw_agent_content.argtypes = [vp, i32]
        L.ssegw_agent_content.restype = A.Bytes
        L.ssegw_agent_has_tool_calls.argtypes = [vp, i32]
        L.ssegw_agent_terminated.argtypes = [vp, i32, C.POINTER(i32)]
        L.ssegw_agent_tool_calls.argtypes = [vp, i32, C.POINTER(A.ToolCall), C.c_size_t]
        L.ssegw_agent_tool_calls.restype = C.c_size_t
        st = i32()
        self.g = L.ssegw_new(device, max_conns, bytes_per_batch, C.byref(st))
        if not self.g:
            raise A.SseError(st.value, "ssegw_new")
        self._buf = C.create_string_buffer(1 << 20)

    def close(self):
        if self.g:
            self.L.ssegw_free(self.g)
            self.g = None

    def stream_chat_completions(self, mode: int) -> int:
        sid = self.L.ssegw_stream_chat_completions(self.g, mode)
        if sid < 0:
            raise RuntimeError("no free connection slot")
        return sid

    def upstream_write(self, sid: int, data: bytes) -> int:
        return self.L.ssegw_upstream_write(self.g, sid, data, len(data))

    def upstream_close(self, sid: int):
        self.L.ssegw_upstream_close(self.g, sid)

    def pump(self) -> int:
        rc = self.L.ssegw_pump(self.g)
        if rc < 0:
            raise A.SseError(rc, "ssegw_pump")
        return rc

    def _recv(self, fn, sid):
        n = C.c_size_t()
        rc = fn(self.g, sid, self._buf, len(self._buf), C.byref(n))
        if rc == 1:
            return self._buf.raw[:n.value]
        if rc == 0:
            return None
        if rc == -1:
            raise EOFError
        raise A.SseError(rc, "recv")

    def recv(self, sid: int):
        """One channel element, None if nothing is available yet; raises EOFError when the channel is closed."""
        return self._recv(self.L.ssegw_recv, sid)

    def proxy_stream(self) -> int:
        """handleStreamingRequest (api/routes.go:129-232): the raw /proxy stream."""
        sid = self.L.ssegw_proxy_stream(self.g)
        if sid < 0:
            raise RuntimeError("no free connection slot")
        return sid

    def proxy_step(self, sid: int):
        """One turn of the c.Stream callback: the line to write, None if none has arrived, EOFError when the loop ends."""
        return self._recv(self.L.ssegw_proxy_step, sid)

    def agent_recv(self, sid: int):
        return self._recv(self.L.ssegw_agent_recv, sid)

    def agent_state(self, sid: int):
        b = self.L.ssegw_agent_content(self.g, sid)
        content = C.string_at(b.p, b.n) if b.n else b""
        fin = C.c_int()
        term = bool(self.L.ssegw_agent_terminated(self.g, sid, C.byref(fin)))
        arr = (A.ToolCall * 64)()
        n = self.L.ssegw_agent_tool_calls(self.g, sid, arr, 64)
        g = lambda x: C.string_at(x.p, x.n) if x.n else b""
        calls = [dict(id=g(arr[i].id), type=g(arr[i].type), name=g(arr[i].name), args=g(arr[i].arguments)) for i in range(min(n, 64))]
        return content, bool(self.L.ssegw_agent_has_tool_calls(self.g, sid)), term, fin.value, calls

    def release(self, sid: int):
        self.L.ssegw_release_stream(self.g, sid)

    def mcp_write_loop(self, sid: int):
        """handleMCPStreamingRequest's writer (api/middlewares/mcp.go:253-299) over the agent channel of one stream:
        returns (bytes written to the client so far, status_503, ended). Call after pump() until ended."""
        out = bytearray()
        status_503 = False
        while True:
            try:
                fr = self.agent_recv(sid)
            except EOFError:                     # channel closed without the terminal frame (mcp.go:256-259)
                return bytes(out), status_503, True
            if fr is None:
                return bytes(out), status_503, False
            flag = C.c_int(0)
            stop = self.L.ssegw_mcp_writer_step(fr, len(fr), C.byref(flag))
            status_503 = status_503 or bool(flag.value)
            out += fr
            if stop:
                return bytes(out), status_503, True
