// Internal declarations shared by the host-side translation units of libdogeray_b200.
#pragma once
#include "dogeray_b200.h"
#include <cstdarg>
#include <memory>
#include <string>
#include <utility>
#include <vector>

// std::allocator whose value-less construct() default-initialises: resize(n) on a vector of PODs then leaves the
// memory untouched, so the parse threads are the first to touch (and fill) the pages they own
template <class T> struct drb_default_init_alloc : std::allocator<T> {
    template <class U> struct rebind { using other = drb_default_init_alloc<U>; };
    drb_default_init_alloc() = default;
    template <class U> drb_default_init_alloc(const drb_default_init_alloc<U>&) noexcept {}
    template <class U> void construct(U* p) noexcept { ::new ((void*)p) U; }
    template <class U, class... A> void construct(U* p, A&&... a) { ::new ((void*)p) U(std::forward<A>(a)...); }
};
using drb_object_vector = std::vector<drb_object, drb_default_init_alloc<drb_object>>;

struct drb_host_scene {
    drb_settings settings;
    drb_object_vector objects;
    std::vector<std::string> tex_paths;   // candidate texture files, sorted
    int64_t skipped = 0;                  // lines that were not turned into objects
    std::string first_warning;
    mutable bool pinned = false;          // objects[] page-locked by the first drb_scene_create (scene.cu)
    mutable int64_t renderable = -1;      // objects that go into the tree, counted on first use (drb_host_scene_num_renderable)
    mutable std::vector<char> tex_used;   // per texture: some object or the settings line names it (same pass as `renderable`)
};
// one parallel pass over the objects that fills `renderable` and `tex_used` (idempotent; not thread-safe per scene)
void drb_host_scene_summarise(const drb_host_scene* hs);

// the rule for "this object line becomes a primitive": SURVEY.md App. B.9 (types other than 0 / 2 are undefined
// behaviour in the reference) and junk lines that stop before the geometry columns; ncols == 0 = made in memory
#ifdef __CUDACC__
__host__ __device__
#endif
inline bool drb_object_renderable(const drb_object& o)
{
    if (o.type == 2) return o.ncols == 0 || o.ncols >= 16;
    if (o.type == 0) return o.ncols == 0 || o.ncols >= 10;
    return false;
}
// drb_host_scene_num_renderable (public) counts them once, in parallel, and remembers: the device build sizes its
// arrays from it without reading a count back from the GPU
// releases the page lock, if any (defined in scene.cu, the only TU that talks to CUDA about it)
void drb_host_scene_unpin(drb_host_scene* hs);

// thread-local last-error string behind drb_last_error()
void drb_set_error(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
void drb_clear_error();

// texture name lookup of the reference (kernel.cu:1172-1183): first path whose lower-cased
// text contains `query` (query itself is not lower-cased); -1 if none
int drb_find_texture(const std::vector<std::string>& tex_paths, const std::string& query);
std::vector<std::string> drb_scan_textures(const char* tex_dir);

// decoded texture (RGBA8, alpha 0, rows top-down)
struct drb_image {
    int w = 0, h = 0;
    std::vector<uint8_t> rgba;
};
int drb_load_ppm(const std::string& path, drb_image& out);
