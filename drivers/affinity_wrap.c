/*
 * Link-time shim for the reference's speedDecode / speedEncode drivers, which
 * pin their worker thread to CPU 16 (speedDecode/speedDecode.c:23,145-148) and
 * die with EINVAL on machines with fewer CPUs.  Linked with
 * -Wl,--wrap=pthread_attr_setaffinity_np, so the driver source stays unchanged:
 * a CPU set with no online CPU in it is replaced by "any CPU".
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <sched.h>
#include <unistd.h>

int __real_pthread_attr_setaffinity_np(pthread_attr_t *attr, size_t cpusetsize, const cpu_set_t *cpuset);

int __wrap_pthread_attr_setaffinity_np(pthread_attr_t *attr, size_t cpusetsize, const cpu_set_t *cpuset)
{
    cpu_set_t allowed;
    if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) {
        for (int cpu = 0; cpu < CPU_SETSIZE; cpu++)
            if (CPU_ISSET(cpu, cpuset) && CPU_ISSET(cpu, &allowed))
                return __real_pthread_attr_setaffinity_np(attr, cpusetsize, cpuset);
        return 0; /* requested CPU does not exist here: leave the thread unpinned */
    }
    return __real_pthread_attr_setaffinity_np(attr, cpusetsize, cpuset);
}
