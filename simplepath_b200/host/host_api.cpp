// C entry points of the host-side plugin library (libsphost.so) for callers that are not C++: parse a .sp file with
// the reference's own FileParser, flatten the resulting sp::Scene (scene_flattener.cpp) and hand out the POD view of
// include/spcu.h.  bench.py and the Python mirror (simplepath_b200/host.py) go through these; the C++ driver uses
// sp::CudaIntegrator directly.  Built against the reference's headers and objects (see Makefile); this library runs the
// reference's scene I/O exactly as its executable would and contains none of its rendering loop.
#include "flat_scene.h"

#include "base/FileParser.h"
#include "base/Logger.h"
#include "base/Scene.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <memory>

namespace sp {
// defined in the reference's main.cpp (:33), which this library does not link
__attribute__((weak)) int k_pretty_print_key = -1;
} // namespace sp

namespace {

using spb200::arm_exit_guard;

void set_error(char* err, size_t errlen, const char* what)
{
    if (err && errlen) {
        std::snprintf(err, errlen, "%s", what);
    }
}

} // namespace

struct sphost_scene
{
    std::unique_ptr<sp::Scene> scene;
    spb200::FlatScene          flat;
};

extern "C" {

// Parse + flatten.  Relative asset paths inside the file resolve against the file's directory.
sphost_scene* sphost_load(const char* sp_path, char* err, size_t errlen)
{
    namespace fs = std::filesystem;
    try {
        sp::Logger::set_level(sp::Logger::LoggingLevel::error);
        const fs::path path = fs::absolute(sp_path);
        std::ifstream  ins(path);
        if (!ins) {
            throw std::runtime_error("cannot open " + path.string());
        }
        const fs::path old = fs::current_path();
        fs::current_path(path.parent_path());
        auto s = std::make_unique<sphost_scene>();
        try {
            s->scene = std::make_unique<sp::Scene>(sp::parse_file(ins));
        } catch (...) {
            fs::current_path(old);
            arm_exit_guard();
            throw;
        }
        fs::current_path(old);
        arm_exit_guard();
        s->flat = spb200::flatten_scene(*s->scene);
        return s.release();
    } catch (const std::exception& e) {
        set_error(err, errlen, e.what());
    } catch (...) {
        set_error(err, errlen, "unknown exception");
    }
    return nullptr;
}

void sphost_free(sphost_scene* s)
{
    delete s;
}

const spcu_flat_scene* sphost_flat(const sphost_scene* s)
{
    return s ? &s->flat.view : nullptr;
}

const char* sphost_output_file_name(const sphost_scene* s)
{
    return s ? s->scene->output_file_name.c_str() : "";
}

// spp x 2 floats: RSequenceSampler::get_next_2D for samples 0..spp-1 (the same for every pixel, see jitter_table()).
void sphost_jitter(unsigned spp, float* out)
{
    const auto t = spb200::jitter_table(spp);
    std::memcpy(out, t.data(), t.size() * sizeof(float));
}

} // extern "C"
