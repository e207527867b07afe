"""`real`/`sys`/`user` time of a call, for meta/all.meta (role of chiron/utils/unix_time.py:11-26)."""
import resource
import time


def unix_time(function, args=tuple(), kwargs=None):
    kwargs = kwargs or {}
    t0, r0 = time.time(), resource.getrusage(resource.RUSAGE_SELF)
    function(*args, **kwargs)
    r1, t1 = resource.getrusage(resource.RUSAGE_SELF), time.time()
    return {"real": t1 - t0, "sys": r1.ru_stime - r0.ru_stime, "user": r1.ru_utime - r0.ru_utime}
