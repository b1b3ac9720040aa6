// HARNESS (oracle/_ref and shim/_build builds only): see mat4x4.hpp.
#pragma once
namespace glm {
struct vec2 {
    float x, y;
};
}  // namespace glm
