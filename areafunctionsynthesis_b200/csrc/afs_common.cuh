// afs_common.cuh -- constants and device helpers shared by the TL and TDS kernels.
//
// Physical constants: reference Constants.h:8-16 (CGS units, Flanagan 1965).
// Section layout: reference Tube.h:53-79.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/afs_b200.h"

namespace afs {

constexpr double kRho = 1.14e-3;      // AMBIENT_DENSITY_CGS
constexpr double kC = 3.5e4;          // SOUND_VELOCITY_CGS
constexpr double kMu = 1.86e-4;       // AIR_VISCOSITY_CGS
constexpr double kKappa = 1.4;        // ADIABATIC_CONSTANT
constexpr double kSpecHeat = 0.24;    // SPECIFIC_HEAT_CGS
constexpr double kHeatCond = 0.055e-3;  // HEAT_CONDUCTION_CGS
constexpr double kMinArea = 0.1e-2;   // Tube.cpp:12 / TlModel.cpp:11 / TdsModel.cpp:19
constexpr double kPi = 3.14159265358979323846;

constexpr int NS = AFS_NUM_SECTIONS;  // 93
constexpr int NC = AFS_NUM_CURRENTS;  // 97
constexpr int TRACHEA0 = 0, TRACHEA_LAST = 22, GLOT_LO = 23, GLOT_UP = 24, PHAR0 = 25, PHAR_LAST = 40;
constexpr int MOUTH0 = 41, MOUTH_LAST = 64, NOSE0 = 65, NOSE_LAST = 83, FOSSA0 = 84, FOSSA_LAST = 88;
constexpr int SINUS0 = 89, SINUS_LAST = 92, FOSSA_COUPLING = 3;

// Geometry and wall data of one static (non pharynx/mouth/glottis) section of Tube::Tube() (Tube.cpp:79-260,
// default lengths Tube.cpp:29-32); the velum port (nose sections 0-3) is handled by the callers.
struct StaticSection {
  double area, length, volume, Mw, Bw, Kw;
};

__host__ __device__ inline double nose_area(int i) {  // Tube.cpp:155-157
  static constexpr double t[19] = {1.63, 2.07, 2.72, 3.59, 4.24, 3.26, 3.04, 3.04, 2.72, 2.5,
                            2.39, 2.39, 1.85, 0.76, 1.41, 1.74, 1.30, 1.74, 0.76};
  return t[i];
}

__host__ __device__ inline StaticSection static_section(int s) {
  static constexpr double sinus_vol[4] = {11.3, 6.8, 33.0, 6.2};      // Tube.cpp:159-161
  static constexpr double neck_len[4] = {0.3, 0.3, 0.45, 1.0};
  static constexpr double neck_area[4] = {0.185, 0.185, 0.145, 0.11};
  StaticSection r;
  r.Mw = 2.1; r.Bw = 800.0; r.Kw = 84500.0;  // Tube.cpp:8-10
  if (s <= TRACHEA_LAST) {                  // Tube.cpp:94-119
    r.area = (s == 0) ? 4.0 : (s == 1) ? 3.0 : 2.5;
    r.length = 23.0 / 23.0;
    r.volume = r.area * r.length;
    r.Mw = 0.25; r.Bw = 1000.0;
  } else if (s >= NOSE0 && s <= NOSE_LAST) {  // Tube.cpp:164-187
    r.area = nose_area(s - NOSE0);
    r.length = (11.4 / 11.4) * 0.6;
    r.volume = r.area * r.length;
  } else if (s >= FOSSA0 && s <= FOSSA_LAST) {  // Tube.cpp:240-258
    const int i = s - FOSSA0;
    r.area = (2.0 * 2.0 / 3.0) * (1.0 - (i + 0.5) / 5.0);
    r.length = 3.0 / 5.0;
    r.volume = r.area * r.length;
  } else {  // sinuses, Tube.cpp:193-214
    const int i = s - SINUS0;
    r.area = neck_area[i]; r.length = neck_len[i]; r.volume = sinus_vol[i];
    r.Mw = 0.0; r.Bw = 6500.0;
  }
  return r;
}

// Velum port: Tube::setVelumOpening (Tube.cpp:402-416), nose section i < 4.
__host__ __device__ inline double velum_port_area(double opening, int i) {
  double a = opening + ((double)(i * i) * (4.24 - opening)) / (double)(4 * 4);
  return a < kMinArea ? kMinArea : a;
}

}  // namespace afs
