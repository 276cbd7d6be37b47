/*
 * Scene recipes.
 *  (1) The reference's golden-image scenes, restated from src/bin/unittests/RenderTests.cpp (one function per
 *      TEST_F, cited below), so the parity tests read like the reference's own tests.
 *  (2) The seeded synthetic workloads of SURVEY.md §8(d) (C1 Cornell, C2 Sponza-class atrium, C3 fog,
 *      C4 instanced forest, C5 progressive emissive spheres) used by bench.py and the offlinerender driver.
 */
#pragma once
#include <string>
#include <vector>
#include "vengine.hpp"

namespace vengine {
namespace scenes {

struct Options {
    int textureSize = 1024; /* procedural texture edge for the synthetic scenes */
    float scale = 1.0f;     /* geometry detail multiplier for the synthetic scenes (1 = BASELINE size) */
    int camera = 0;         /* C5: 0 perspective, 1 orthographic */
};

std::vector<std::string> list();
/* Clears the engine's scene, builds `name`, calls scene.update() and fills renderInfo() with the recipe's
 * settings (resolution, samples, batch size, depth, file name). Returns false for an unknown name. */
bool build(Engine &engine, const std::string &name, const Options &opt = Options());
/* the render sequence of the reference's BallOnPlane demo (src/bin/offlinerender/PtSceneBallOnPlane.cpp:44-55): moves the
 * camera of an already built "BallOnPlane" scene to frame 0..7 of its orbit and names the output file after the frame */
constexpr int kBallOnPlaneFrames = 8;
void ballOnPlaneFrame(Engine &engine, int frame);

}  // namespace scenes
}  // namespace vengine
