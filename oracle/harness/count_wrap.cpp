// Link-time ray counter for the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY; our own code).
//
// The reference cannot count the rays it traces.  Binaries linked with
//   -Wl,--wrap=_ZN10BVHMethods16FindIntersectionERK3RayRSt6vectorIP5ShapeSaIS5_EERS3_IP8InstanceSaISA_EE
// route every call of BVHMethods::FindIntersection (src/Helper.cpp:18-80: THE closest-hit query - primary, reflection,
// refraction and every light's IsShadow go through it) made from another object file to __wrap_..., which counts and
// forwards to the real function.  No reference source is touched; the reference objects are the same ones ref_raytracer
// links.  Counts are kept per thread and folded into the total when a worker thread exits (the reference joins its 8
// workers before the frame is saved, src/Scene.cpp:349-356).
#include "Helper.h"
#include <atomic>

extern "C" ReturnVal __real__ZN10BVHMethods16FindIntersectionERK3RayRSt6vectorIP5ShapeSaIS5_EERS3_IP8InstanceSaISA_EE(
    const Ray& ray, std::vector<Shape*>& objects, std::vector<Instance*>& instances);

namespace {
std::atomic<unsigned long long> g_total{0};
struct Local {
    unsigned long long n = 0;
    ~Local() { g_total.fetch_add(n, std::memory_order_relaxed); }
};
thread_local Local t_local;
}

extern "C" ReturnVal __wrap__ZN10BVHMethods16FindIntersectionERK3RayRSt6vectorIP5ShapeSaIS5_EERS3_IP8InstanceSaISA_EE(
    const Ray& ray, std::vector<Shape*>& objects, std::vector<Instance*>& instances)
{
    t_local.n++;
    return __real__ZN10BVHMethods16FindIntersectionERK3RayRSt6vectorIP5ShapeSaIS5_EERS3_IP8InstanceSaISA_EE(ray, objects, instances);
}

// Queries counted so far by threads that have exited, plus the calling thread's own.
extern "C" unsigned long long rt795_ray_queries(int reset)
{
    unsigned long long v = g_total.load() + t_local.n;
    if (reset) { g_total.store(0); t_local.n = 0; }
    return v;
}
