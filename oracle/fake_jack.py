"""TEST INFRASTRUCTURE ONLY -- stand-in for the ``jack`` (JACK-Client) module.

The reference's ``JackClient`` (``_jack_client.py:239-289``) talks to a JACK server through this
module; there is no JACK here.  The stand-in gives ``jack.Client`` the attributes the reference
reads while it builds a client (block size, sample rate, status flags, port lists) and swallows
every other call, so that ``JackRenderer`` / ``JackRendererBenchmark`` objects can be constructed
in-process and their ``_process`` callback driven directly by a test.
"""
import types
from unittest.mock import MagicMock

DEFAULT_SAMPLERATE = 48000
DEFAULT_BLOCKSIZE = 256


class _Ports(list):
    def register(self, name):
        port = MagicMock(name=f"port:{name}")
        self.append(port)
        return port

    def clear(self):
        del self[:]


class Client:
    def __init__(self, name="client", *args, **kwargs):
        self.name = name
        self.blocksize = DEFAULT_BLOCKSIZE
        self.samplerate = DEFAULT_SAMPLERATE
        self.status = types.SimpleNamespace(server_started=False, name_not_unique=False)
        self.inports, self.outports = _Ports(), _Ports()

    def __getattr__(self, item):  # callbacks setters, activate / deactivate / close, cpu_load, ...
        return MagicMock(name=f"jack.Client.{item}")


def client_name_size():
    return 64


def build_module():
    mod = types.ModuleType("jack")
    mod.Client = Client
    mod.client_name_size = client_name_size
    mod.JackError = type("JackError", (Exception,), {})
    mod.CallbackExit = type("CallbackExit", (Exception,), {})
    mod.set_error_function = lambda f: f
    mod.set_info_function = lambda f: f
    mod.__getattr__ = lambda item: MagicMock(name=f"jack.{item}")
    return mod
