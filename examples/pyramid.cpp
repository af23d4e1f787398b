// Headless C++ port of the reference's demo5 (examples/samples/src/main.rs:157-178) on the B200 backend.
// build: g++ -std=c++17 -Iinclude examples/pyramid.cpp -Lsylt2d_b200 -lsylt2d_b200 -Wl,-rpath,$PWD/sylt2d_b200 -o examples/pyramid
#include <cstdio>
#include "sylt2d.hpp"
using namespace sylt2d;
int main(int argc, char** argv) {
    int steps = argc > 1 ? atoi(argv[1]) : 300;
    try {
        World world(Vec2(0.0f, -10.0f), 100);   // ITERATIONS = 100 (main.rs:10)
        world.world_context.warm_starting = true;
        Body ground = Body::make(Vec2(100.0f, 20.0f), FLT_MAX);
        ground.friction = 0.2f;
        ground.position = Vec2(0.0f, -0.5f * ground.width.y);
        world.add_body(ground);
        Vec2 x(-6.0f, 0.75f);
        for (int i = 0; i < 12; i++) {
            Vec2 y = x;
            for (int j = i; j < 12; j++) {
                Body b = Body::make(Vec2(1.0f, 1.0f), 10.0f);
                b.friction = 0.2f;
                b.position = y;
                world.add_body(b);
                y.x += 1.125f;
            }
            x.x += 0.5625f;
            x.y += 2.0f;
        }
        for (int s = 0; s < steps; s++) world.step(1.0f / 60.0f);
        auto b = world.bodies();
        float ymax = -1e9f;
        for (auto& s : b) ymax = s.y > ymax ? s.y : ymax;
        printf("bodies %zu arbiters %zu top y %.4f\n", b.size(), world.arbiters().size(), ymax);
        return (ymax > 11.0f && ymax < 12.0f) ? 0 : 1;
    } catch (const Error& e) {
        fprintf(stderr, "error %d: %s\n", e.code, e.what());
        return 2;
    }
}
