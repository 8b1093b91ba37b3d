"""pytest plugin (-p asyncio_shim): runs `async def` tests with asyncio.run — pytest-asyncio is not in this image.
No support for async fixtures (the reference's hot-path unit tests use none)."""
import asyncio
import inspect

import pytest


def pytest_configure(config):
    config.addinivalue_line("markers", "asyncio: run the coroutine test with asyncio.run")


@pytest.hookimpl(tryfirst=True)
def pytest_pyfunc_call(pyfuncitem):
    if inspect.iscoroutinefunction(pyfuncitem.obj):
        kw = {a: pyfuncitem.funcargs[a] for a in pyfuncitem._fixtureinfo.argnames}
        asyncio.run(pyfuncitem.obj(**kw))
        return True
