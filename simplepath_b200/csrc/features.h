// Compile-time scene feature sets.  Every shading / traversal function is a template over one of these; a feature that
// is `false` removes its code (and its registers) from the kernel instead of leaving a never-taken branch behind.  The
// host picks the instantiation that matches the uploaded scene (spcu_api.cu: scene_features).  Why it matters: the
// shading kernels were limited by instruction fetch and by registers that only the microfacet / image-based-light code
// needs (ncu: stall_no_instruction up to 7.8 cycles per issued instruction, 96-125 registers), even on scenes that
// contain neither.
#pragma once

namespace spcu {

struct FeatFull // anything the flattener can emit
{
    static constexpr bool bvh           = true; // the geometry accelerator has internal nodes
    static constexpr bool triangles     = true;
    static constexpr bool microfacet    = true; // MicrofacetReflection / Beckmann (and with it the 16-sample albedo estimate)
    static constexpr bool specular_bxdf = true; // SpecularReflectionBRDF inside a OneSampleMaterial
    static constexpr bool ibl           = true; // ImageBasedEnvironmentLight
    static constexpr bool multi_bxdf    = true; // a OneSampleMaterial may hold several BxDFs (selection weights, one-sample MIS)
    static constexpr bool nested_coats  = true; // a ClearcoatMaterial's base may be another ClearcoatMaterial
    static constexpr int  id            = 0;
};

// spheres and planes only; every OneSampleMaterial is ONE Lambert BxDF whose albedo has a normal (finite, non-zero)
// luminance — its selection weight x / x is then exactly 1 and the one-sample combination is the BxDF itself
// (materials/Material.h:545-572, 669-715) —, with or without ONE clearcoat over it; sphere and constant environment lights
struct FeatAnalytic
{
    static constexpr bool bvh           = false;
    static constexpr bool triangles     = false;
    static constexpr bool microfacet    = false;
    static constexpr bool specular_bxdf = false;
    static constexpr bool ibl           = false;
    static constexpr bool multi_bxdf    = false;
    static constexpr bool nested_coats  = false;
    static constexpr int  id            = 1;
};

constexpr int kNumFeatureSets = 2;

} // namespace spcu
