#!/usr/bin/env python3
"""Applies the drop-in hunks of INTEGRATION.md to a build-time view of the reference (never to /root/reference, never
into git): after mkfarm.sh has made <farm> a tree of symlinks into <ref>, this replaces three of them by patched files.

  Integrators/Integrator.h    enum IntegratorType gains `Cuda`                               (:18-28)
  Integrators/Integrator.cpp  string_to_integrator_type accepts "cuda[_<inner>]"             (:25-51)
  main_dropin.cpp             = main.cpp with: #include, factory case (:36-49), render() dispatch (:109-130)
  base/Scene.h                create_acceleration_structure (:27-45) can leave the GEOMETRY accelerator unbuilt — the
                              top-level list [unbounded..., bounded in pre-construction order] without the BVH over the
                              bounded part — when sp::CudaIntegrator was selected on the command line: it builds that BVH
                              on the device (spcu_upload_scene_build), bit-identical to BVHAccelerator(first, part_it)

Each hunk is an insertion next to an anchor line that must occur exactly once; a reference that has moved on fails the
build loudly instead of silently compiling something else."""
import sys
from pathlib import Path


def patch(text: str, anchor: str, insert: str, before: bool, what: str) -> str:
    if text.count(anchor) != 1:
        raise SystemExit(f"apply_dropin: anchor for '{what}' occurs {text.count(anchor)} times (expected 1): {anchor!r}")
    return text.replace(anchor, insert + anchor if before else anchor + insert)


def main() -> None:
    ref, farm = Path(sys.argv[1]), Path(sys.argv[2])

    h = (ref / "Integrators/Integrator.h").read_text()
    h = patch(h, "    Whitted\n};", "", False, "enum end")  # presence check
    h = h.replace("    Whitted\n};", "    Whitted,\n    Cuda\n};")
    out = farm / "Integrators/Integrator.h"
    out.unlink(missing_ok=True)
    out.write_text(h)

    c = (ref / "Integrators/Integrator.cpp").read_text()
    c = patch(c, '    throw std::runtime_error("Unknown integrator type");',
              "    if (sp::CudaIntegrator::select(s)) {\n        return IntegratorType::Cuda;\n    }\n\n", True, "string_to_integrator_type")
    c = patch(c, '#include "Integrator.h"\n', '#include "cuda_integrator.h"\n', False, "include")
    out = farm / "Integrators/Integrator.cpp"
    out.unlink(missing_ok=True)
    out.write_text(c)

    s = (ref / "base/Scene.h").read_text()
    s = patch(s, "template <typename Iterator>\n    requires std::random_access_iterator<Iterator>\nListAccelerator create_acceleration_structure(",
              "// set by sp::CudaIntegrator::select() before the scene is parsed (simplepath_b200/host/cuda_integrator.cpp)\n"
              "inline bool g_defer_geometry_accelerator = false;\n\n", True, "defer flag")
    s = patch(s, "    const auto      bounded_accelerator = std::make_shared<BVHAccelerator>(first, part_it);\n",
              "    if constexpr (std::is_convertible_v<typename std::iterator_traits<Iterator>::value_type,\n"
              "                                        std::shared_ptr<const GeometricPrimitive>>) {\n"
              "        if (g_defer_geometry_accelerator) {\n"
              "            // the order std::partition left behind IS the input of BVHAccelerator(first, part_it): kept as a plain list\n"
              "            ListAccelerator unbuilt(part_it, last);\n"
              "            for (auto it = first; it != part_it; ++it) {\n"
              "                unbuilt.push_back(*it);\n"
              "            }\n"
              "            unbuilt.shrink_to_fit();\n"
              "            return unbuilt;\n"
              "        }\n"
              "    }\n", True, "deferred geometry accelerator")
    out = farm / "base/Scene.h"
    out.unlink(missing_ok=True)
    out.write_text(s)

    m = (ref / "main.cpp").read_text()
    m = patch(m, "namespace fs = std::filesystem;", '#include "cuda_integrator.h"\n\n', True, "main include")
    m = patch(m, "    case sp::IntegratorType::Whitted: return std::make_unique<sp::WhittedIntegrator>();\n",
              "    case sp::IntegratorType::Cuda: return std::make_unique<sp::CudaIntegrator>();\n", False, "factory case")
    m = patch(m, "    for (int i = 0; i < num_threads; ++i) {\n        threads.emplace_back(&render_thread,",
              "    // whole-frame integrators render here; the per-pixel thread loop below is skipped for them\n"
              "    if (const auto* cuda = dynamic_cast<const sp::CudaIntegrator*>(&integrator);\n"
              "        cuda && cuda->render_frame(scene, num_pixel_samples, image)) {\n"
              "        num_threads = 0;\n"
              "    }\n\n", True, "render dispatch")
    (farm / "main_dropin.cpp").write_text(m)


if __name__ == "__main__":
    main()
