// Owning storage for a spcu_flat_scene (include/spcu.h) plus the entry points of the flattener.
// Host side of the drop-in boundary; compiled against the reference's headers.
#pragma once

#include "spcu.h"

#include <string>
#include <vector>

namespace sp {
class Scene;
}

namespace spb200 {

struct FlatScene
{
    spcu_flat_scene view{}; // pointers below are wired into this by finalize()

    std::vector<spcu_bvh_node>   geom_nodes;
    std::vector<spcu_prim_geom>  geom_prims;
    std::vector<spcu_prim_shade> geom_shade;
    std::vector<uint32_t>        geom_meta;

    std::vector<spcu_bvh_node> light_nodes;
    std::vector<spcu_light>    lights;
    std::vector<uint32_t>      light_order;

    std::vector<spcu_material> materials;
    std::vector<spcu_bxdf>     bxdfs;
    std::vector<float>         float_pool;

    // The geometry accelerator was left unbuilt by the patched create_acceleration_structure (apply_dropin.py): geom_prims
    // hold [unbounded..., bounded in pre-construction order], view.geom only n_prims / n_unbounded, geom_bounds the reference's
    // own world bounds of the bounded primitives (spheres need them; triangle bounds are recomputed on the device).
    bool                     geom_unbuilt = false;
    std::vector<spcu_bounds> geom_bounds;

    void finalize();
};

// Walks the reference's own object graph (private members; this TU is built with -fno-access-control)
// and emits the POD scene.  Throws std::runtime_error on anything the device path cannot represent.
FlatScene flatten_scene(const sp::Scene& scene);

// spp x 2 pixel jitter exactly as main.cpp:67-71,96 draws it (RSequenceSampler::get_next_2D).
std::vector<float> jitter_table(unsigned spp);

// The reference's AccumulatedLogger singleton joins its worker thread from a static destructor and never returns
// (base/AccumulatedLogger.h:34-37 starts the thread before the condition variable it waits on is constructed), so a
// process that parsed a scene hangs at exit.  This registers (once) an exit handler that flushes stdio and leaves
// through _exit with the real status; called after the logger exists, it runs before the logger's destructor.
void arm_exit_guard();

// Binary (de)serialisation so tests on a box without the reference can reuse a flattened scene.
void      save_flat_scene(const FlatScene& fs, const std::string& path);
FlatScene load_flat_scene(const std::string& path);

} // namespace spb200
