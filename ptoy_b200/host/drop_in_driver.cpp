// A caller written ONLY against the reference's public `Particles` interface (particles.h),
// doing what main.cpp / ptoy_wasm.cpp / control.cpp do: per frame
// SetRendererReadyForNext(true); step(GetTime() + dt); read the getters — plus a scripted UI
// session (gravity, point force, bonds, freeze, portals, resize).  The SAME source is linked once
// against the reference's particles.cpp (test infrastructure, see the Makefile) and once against
// the drop-in facade + libptoy_b200.so (drop_in_driver_b200); tests/test_dropin_gpu.py compares
// their outputs.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "particles.h"

static void dump(const char* tag, Particles& ps) {
  const auto& bd = ps.GetBlockData();
  const auto bbi = ps.GetBlockById();
  std::printf("@%s t=%.9g steps=%zu n=%zu buf=%zu bonds=%zu frozen=%zu norender=%zu portals=%zu stage=%d grav=%d\n", tag,
              (double)ps.GetTime(), ps.GetNumSteps(), ps.GetNumParticles(), ps.GetParticles().size(),
              ps.GetBonds().size(), ps.GetFrozen().size(), ps.GetNoRendering().size(), ps.GetPortals().size(),
              ps.portal_stage_, (int)ps.GetGravity());
  const RectVect d = ps.GetDomain();
  std::printf("#domain %.9g %.9g %.9g %.9g\n", (double)d.A.x, (double)d.A.y, (double)d.B.x, (double)d.B.y);
  for (auto& pair : ps.GetPortals())
    for (auto& p : pair)
      std::printf("#portal %.9g %.9g %.9g %.9g blocks=%zu\n", (double)p.begin.x, (double)p.begin.y, (double)p.end.x,
                  (double)p.end.y, p.blocks.size());
  for (auto b : ps.GetBonds()) std::printf("#bond %d %d\n", b.first, b.second);
  for (int f : ps.GetFrozen()) std::printf("#frozen %d\n", f);
  for (size_t id = 0; id < bbi.size(); ++id) {
    if (bbi[id].first == blocks::kBlockNone) continue;
    const Vect p = bd.position[bbi[id].first][bbi[id].second];
    std::printf("p %zu %zu %.9g %.9g\n", id, bbi[id].first, (double)p.x, (double)p.y);
  }
}

int main(int argc, char** argv) {
  const int frames = argc > 1 ? std::atoi(argv[1]) : 3;
  Particles ps;
  dump("init", ps);
  const Scal dt = 0.01;  // 20 sub-steps per frame
  auto frame = [&]() {
    ps.SetRendererReadyForNext(true);
    ps.step(ps.GetTime() + dt, false);
  };
  for (int f = 0; f < frames; ++f) frame();
  dump("free", ps);
  // UI session on the current snapshot
  auto pos_of = [&](size_t id) {  // by id: the in-block order of the snapshot is not part of the contract
    const auto bbi = ps.GetBlockById();
    return ps.GetBlockData().position[bbi[id].first][bbi[id].second];
  };
  const Vect a = pos_of(10), b = pos_of(11), c = pos_of(36), d = pos_of(300);
  ps.BondsStart(a); ps.BondsMove(b); ps.BondsMove(c); ps.BondsStop(c);
  ps.FreezeStart(d); ps.FreezeStop(d);
  ps.PortalStart(Vect(0.0, 0.2)); ps.PortalMove(Vect(0.0, 0.5)); ps.PortalStop(Vect(0.0, 0.7));
  ps.PortalStart(Vect(0.3, 0.2)); ps.PortalStop(Vect(0.4, 0.9));
  ps.SetForce(Vect(0.2, -0.5), true);
  ps.SetForceAttractive(true);
  ps.SetGravityVect(Vect(2, -9));
  ps.PushResize(RectVect(Vect(-1.1, -1.05), Vect(1.2, 1.0)));
  for (int f = 0; f < frames; ++f) frame();
  dump("ui", ps);
  ps.RemoveLastPortal();
  ps.SetForce(false);
  ps.SetGravity(false);
  frame();
  dump("end", ps);
  return 0;
}
