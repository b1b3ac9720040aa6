// HARNESS (oracle/_ref and shim/_build builds only): the two glm types the reference's visualization module headers
// (src/domains/visualization/{lineplot,waterfall}/module_impl.hh) name in their render-uniform structs. glm is an
// un-vendored render dependency; nothing on the compute path touches these members.
#pragma once
namespace glm {
struct mat4 {
    float m[16];
    mat4() : m{} {}
    explicit mat4(float d) : m{} { m[0] = m[5] = m[10] = m[15] = d; }
};
}  // namespace glm
