// GJK + EPA convex-convex narrowphase (placeholder until the hull/cylinder path is restated).
namespace pgo {
struct World; struct Body; struct ContactSink;
static void convex_convex(World&, Body&, Body&, ContactSink&) {}
}
