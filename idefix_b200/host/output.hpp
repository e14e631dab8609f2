// output.hpp -- Output: periodic dump / VTK / analysis (reference: src/output/output.hpp:27-120,
// output.cpp:20-400), dump files (src/output/dump.cpp:812-985 writer, :628-810 reader) and legacy
// binary VTK files (src/output/vtk.cpp:60-600).
//
// File formats are the reference's, so pytools/dump_io.py, pytools/vtk_io.py and ParaView read
// them unchanged:
//  * dump.NNNN.dmp: 128-byte header "Idefix <ver> Dump Data little endian", then fields
//    {char name[16]; int32 type; int32 ndim; int32 dims[ndim]; data} (type 0 double, 2 int32),
//    x1,xl1,xr1 ... xr3, all registered variables in alphabetical order, terminated by "eof";
//  * data.NNNN.vtk: "# vtk DataFile Version 2.0", BINARY, RECTILINEAR_GRID (cartesian) or
//    STRUCTURED_GRID, FIELD data GEOMETRY / PERIODICITY / TIME / native coordinates, CELL_DATA
//    scalars as big-endian float.
#include "idefix.hpp"
#ifndef IDEFIX_B200_HOST_OUTPUT_HPP_
#define IDEFIX_B200_HOST_OUTPUT_HPP_

using AnalysisFunc = void (*)(DataBlock &);
using UserDefVariablesContainer = std::map<std::string, IdefixHostArray3D<real>>;
using UserDefVariablesFunc = void (*)(DataBlock &, UserDefVariablesContainer &);

namespace idfxhost {
inline float bigEndian(float v) {
  uint32_t u;
  std::memcpy(&u, &v, 4);
  u = (u >> 24) | ((u >> 8) & 0xff00u) | ((u << 8) & 0xff0000u) | (u << 24);
  std::memcpy(&v, &u, 4);
  return v;
}
inline int32_t bigEndian(int32_t v) {
  uint32_t u = static_cast<uint32_t>(v);
  u = (u >> 24) | ((u >> 8) & 0xff00u) | ((u << 8) & 0xff0000u) | (u << 24);
  return static_cast<int32_t>(u);
}
inline std::string numbered(const std::string &base, int n, const std::string &ext) {
  std::stringstream ss;
  ss << base << "." << std::setfill('0') << std::setw(4) << n << "." << ext;
  return ss.str();
}
}  // namespace idfxhost

class Output {
 public:
  bool forceNoWrite{false};
  bool vtkEnabled{false}, dumpEnabled{false}, analysisEnabled{false}, userDefVariablesEnabled{false};
  real vtkPeriod{0}, vtkLast{0}, dumpPeriod{0}, dumpLast{0}, analysisPeriod{0}, analysisLast{0};
  int vtkFileNumber{0}, dumpFileNumber{0};
  std::string vtkDir{"./"}, dumpDir{"./"};
  UserDefVariablesContainer userDefVariables;

  Output(Input &input, DataBlock &data) {
    forceNoWrite = input.forceNoWrite;
    if (input.CheckEntry("Output", "vtk") > 0) {
      vtkPeriod = input.Get<real>("Output", "vtk", 0);
      if (vtkPeriod >= 0.0) {   // negative period = disabled
        vtkLast = data.t - vtkPeriod;
        vtkEnabled = true;
      }
    }
    if (input.CheckEntry("Output", "vtk_dir") > 0) vtkDir = input.Get<std::string>("Output", "vtk_dir", 0) + "/";
    if (input.CheckEntry("Output", "dmp") > 0) {
      dumpPeriod = input.Get<real>("Output", "dmp", 0);
      if (dumpPeriod >= 0.0) {
        dumpLast = data.t - dumpPeriod;
        dumpEnabled = true;
      }
    }
    if (input.CheckEntry("Output", "dmp_dir") > 0) dumpDir = input.Get<std::string>("Output", "dmp_dir", 0) + "/";
    if (input.CheckEntry("Output", "analysis") > 0) {
      analysisPeriod = input.Get<real>("Output", "analysis", 0);
      analysisLast = data.t - analysisPeriod;
      analysisEnabled = true;
    }
    if (input.CheckEntry("Output", "uservar") > 0) {
      const int nvars = input.CheckEntry("Output", "uservar");
      for (int v = 0; v < nvars; v++) {
        const std::string name = input.Get<std::string>("Output", "uservar", v);
        userDefVariables[name] = IdefixHostArray3D<real>("userVar-" + name, data.np_tot[KDIR], data.np_tot[JDIR],
                                                         data.np_tot[IDIR]);
      }
      userDefVariablesEnabled = true;
    }
  }

  void EnrollAnalysis(AnalysisFunc f) {
    if (!analysisEnabled)
      IDEFIX_WARNING("You are enrolling an analysis function but analysis are not enabled in the input file,"
                     "so no analysis will be performed");
    analysisFunc = f;
  }
  void EnrollUserDefVariables(UserDefVariablesFunc f) {
    if (!userDefVariablesEnabled)
      IDEFIX_WARNING("You are enrolling a user-defined variables function but the userdef variables are not set "
                     "in the input file, so no user-defined variable will be written");
    userDefVariablesFunc = f;
  }

  // output.cpp:162-300: an output fires when t >= last + period
  int CheckForWrites(DataBlock &data) {
    if (forceNoWrite) return 0;
    int nfiles = 0;
    const auto t0 = std::chrono::steady_clock::now();
    if (vtkEnabled && data.t >= vtkLast + vtkPeriod) {
      vtkLast += vtkPeriod;
      WriteVtk(data);
      nfiles++;
      if (vtkLast + vtkPeriod <= data.t && vtkPeriod > 0.0)
        while (vtkLast <= data.t - vtkPeriod) vtkLast += vtkPeriod;
    }
    if (analysisEnabled && data.t >= analysisLast + analysisPeriod) {
      if (!analysisFunc)
        IDEFIX_ERROR("Cannot perform a user-defined analysis without enrollment of your analysis function");
      analysisLast += analysisPeriod;
      analysisFunc(data);
      nfiles++;
      if (analysisLast + analysisPeriod <= data.t && analysisPeriod > 0.0)
        while (analysisLast <= data.t - analysisPeriod) analysisLast += analysisPeriod;
    }
    if (dumpEnabled && data.t >= dumpLast + dumpPeriod) {
      dumpLast += dumpPeriod;
      WriteDump(data);
      nfiles++;
      if (dumpLast + dumpPeriod <= data.t && dumpPeriod > 0.0)
        while (dumpLast <= data.t - dumpPeriod) dumpLast += dumpPeriod;
    }
    elapsed += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return nfiles;
  }
  // next time at which an output can fire (used to size device-resident batches of cycles)
  bool HasPeriodicOutputs() const { return !forceNoWrite && (vtkEnabled || dumpEnabled || analysisEnabled); }
  void ForceWriteDump(DataBlock &data) {
    if (!forceNoWrite) WriteDump(data);
  }
  void ForceWriteVtk(DataBlock &data) {
    if (!forceNoWrite) WriteVtk(data);
  }
  void ResetTimer() { elapsed = 0.0; }
  double GetTimer() const { return elapsed; }

  // ---- dump files --------------------------------------------------------------------------------------
  void WriteDump(DataBlock &data) {
    const std::string name = dumpDir + idfxhost::numbered("dump", dumpFileNumber, "dmp");
    idfx::cout << "Dump: Write file " << name << "..." << std::flush;
    dumpFileNumber++;   // before the scalars are emitted: dump.N stores N+1 (dump.cpp:839), Dump::Read takes it as is
    const auto t0 = std::chrono::steady_clock::now();
    DataBlockHost d(data);
    d.SyncFromDevice();
    std::FILE *f = std::fopen(name.c_str(), "wb");
    if (!f) IDEFIX_ERROR("Unable to open file " + name);
    char header[128] = {};
    std::snprintf(header, sizeof(header), "Idefix %s Dump Data little endian", IDEFIX_VERSION);
    std::fwrite(header, 1, sizeof(header), f);
    for (int dir = 0; dir < 3; dir++) {
      const int n = data.np_int[dir], g = data.nghost[dir];
      const std::string s = std::to_string(dir + 1);
      WriteField(f, "x" + s, 0, {n}, d.x[dir].data() + g);
      WriteField(f, "xl" + s, 0, {n}, d.xl[dir].data() + g);
      WriteField(f, "xr" + s, 0, {n}, d.xr[dir].data() + g);
    }
    // registered entries in alphabetical order (std::map in the reference)
    std::map<std::string, std::function<void()>> entries;
    const auto &names = data.hydro->VcName;
    for (int n = 0; n < static_cast<int>(names.size()); n++) {
#if MHD == YES
      if (n >= BX1 && n < BX1 + DIMENSIONS) continue;   // evolved through Vs (fluid.hpp:607-622)
#endif
      entries["Vc-" + names[n]] = [&, n]() {
        std::vector<double> buf = Interior(d, [&](int k, int j, int i) { return d.Vc(n, k, j, i); }, 0, 0, 0);
        WriteField(f, "Vc-" + names[n], 0, {data.np_int[0], data.np_int[1], data.np_int[2]}, buf.data());
      };
    }
#if MHD == YES
    for (int n = 0; n < DIMENSIONS; n++) {
      entries["Vs-" + data.hydro->VsName[n]] = [&, n]() {
        const int o[3] = {n == 0, n == 1, n == 2};
        std::vector<double> buf = Interior(d, [&](int k, int j, int i) { return d.Vs(n, k, j, i); }, o[0], o[1], o[2]);
        WriteField(f, "Vs-" + data.hydro->VsName[n], 0,
                   {data.np_int[0] + o[0], data.np_int[1] + o[1], data.np_int[2] + o[2]}, buf.data());
      };
    }
#ifdef EVOLVE_VECTOR_POTENTIAL
    // edge-centred: one more point in every active direction perpendicular to the component (dump.cpp:921-930)
    for (int n = 0; n < static_cast<int>(data.hydro->VeName.size()); n++) {
      entries["Ve-" + data.hydro->VeName[n]] = [&, n]() {
        const int dirn = (DIMENSIONS < 3) ? 2 : n;
        const int o[3] = {dirn != 0 && DIMENSIONS >= 1, dirn != 1 && DIMENSIONS >= 2, dirn != 2 && DIMENSIONS >= 3};
        std::vector<double> buf = Interior(d, [&](int k, int j, int i) { return d.Ve(n, k, j, i); }, o[0], o[1], o[2]);
        WriteField(f, "Ve-" + data.hydro->VeName[n], 0,
                   {data.np_int[0] + o[0], data.np_int[1] + o[1], data.np_int[2] + o[2]}, buf.data());
      };
    }
#endif
#endif
    auto scalarD = [&](const std::string &nm, double v) {
      entries[nm] = [&, nm, v]() { WriteField(f, nm, 0, {1}, &v); };
    };
    auto scalarI = [&](const std::string &nm, int32_t v) {
      entries[nm] = [&, nm, v]() { WriteField(f, nm, 2, {1}, &v); };
    };
    scalarD("time", data.t);
    scalarD("dt", data.dt);
    scalarD("vtkLast", vtkLast);
    scalarD("dumpLast", dumpLast);
    scalarD("analysisLast", analysisLast);
    scalarI("vtkFileNumber", vtkFileNumber);
    scalarI("dumpFileNumber", dumpFileNumber);
    scalarI("geometry", GEOMETRY);
    const int32_t per[3] = {data.mygrid->lbound[0] == periodic, data.mygrid->lbound[1] == periodic,
                            data.mygrid->lbound[2] == periodic};
    entries["periodicity"] = [&]() { WriteField(f, "periodicity", 2, {3}, per); };
    for (auto &e : entries) e.second();
    const double zero = 0.0;
    WriteField(f, "eof", 0, {1}, &zero);
    std::fclose(f);
    idfx::cout << "done in " << std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() << " s."
               << std::endl;
  }

  // Dump::Read (dump.cpp:628-810): restores Vc / Vs / time and the output bookkeeping
  bool RestartFromDump(DataBlock &data, int number) {
    if (number < 0) {   // latest existing dump
      for (int n = 0; n < 10000; n++) {
        std::ifstream probe(dumpDir + idfxhost::numbered("dump", n, "dmp"));
        if (!probe) break;
        number = n;
      }
      if (number < 0) {
        idfx::cout << "Dump: cannot find a valid dump to restart from." << std::endl;
        return false;
      }
    }
    const std::string name = dumpDir + idfxhost::numbered("dump", number, "dmp");
    std::FILE *f = std::fopen(name.c_str(), "rb");
    if (!f) {
      idfx::cout << "Dump: cannot open " << name << std::endl;
      return false;
    }
    idfx::cout << "Dump: Reading " << name << "..." << std::flush;
    char header[128];
    if (std::fread(header, 1, 128, f) != 128) IDEFIX_ERROR("truncated dump file");
    DataBlockHost d(data);
    d.SyncFromDevice();
    while (true) {
      char nm[16];
      int32_t type, ndim, dims[4] = {1, 1, 1, 1};
      if (std::fread(nm, 1, 16, f) != 16) break;
      if (std::fread(&type, 4, 1, f) != 1 || std::fread(&ndim, 4, 1, f) != 1) break;
      if (std::fread(dims, 4, ndim, f) != static_cast<size_t>(ndim)) break;
      const std::string field(nm);
      if (field == "eof") break;
      long long n = 1;
      for (int q = 0; q < ndim; q++) n *= dims[q];
      const size_t sz = (type == 0) ? 8 : (type == 3 ? 1 : 4);
      std::vector<char> raw(n * sz);
      if (std::fread(raw.data(), sz, n, f) != static_cast<size_t>(n)) IDEFIX_ERROR("truncated dump file");
      const double *dv = reinterpret_cast<const double *>(raw.data());
      const int32_t *iv = reinterpret_cast<const int32_t *>(raw.data());
      if (field.rfind("Vc-", 0) == 0 || field.rfind("Vs-", 0) == 0) {
        const bool stag = field[1] == 's';
        const auto &list = stag ? data.hydro->VsName : data.hydro->VcName;
        const auto it = std::find(list.begin(), list.end(), field.substr(3));
        if (it == list.end()) {
          IDEFIX_WARNING("Dump: ignoring unknown field " + field);
          continue;
        }
        const int v = static_cast<int>(it - list.begin());
        const int o[3] = {stag && v == 0, stag && v == 1, stag && v == 2};
        if (dims[0] != data.np_int[0] + o[0] || dims[1] != data.np_int[1] + o[1] || dims[2] != data.np_int[2] + o[2])
          IDEFIX_ERROR("Dump: the grid of the restart file does not match the current grid");
        long long c = 0;
        for (int k = data.beg[2]; k < data.end[2] + o[2]; k++)
          for (int j = data.beg[1]; j < data.end[1] + o[1]; j++)
            for (int i = data.beg[0]; i < data.end[0] + o[0]; i++) {
              if (stag) {
#if MHD == YES
                d.Vs(v, k, j, i) = dv[c++];
#endif
              } else {
                d.Vc(v, k, j, i) = dv[c++];
              }
            }
#ifdef EVOLVE_VECTOR_POTENTIAL
      } else if (field.rfind("Ve-", 0) == 0) {
        const auto &list = data.hydro->VeName;
        const auto it = std::find(list.begin(), list.end(), field.substr(3));
        if (it == list.end()) {
          IDEFIX_WARNING("Dump: ignoring unknown field " + field);
          continue;
        }
        const int v = static_cast<int>(it - list.begin());
        const int dirn = (DIMENSIONS < 3) ? 2 : v;
        const int o[3] = {dirn != 0 && DIMENSIONS >= 1, dirn != 1 && DIMENSIONS >= 2, dirn != 2 && DIMENSIONS >= 3};
        if (dims[0] != data.np_int[0] + o[0] || dims[1] != data.np_int[1] + o[1] || dims[2] != data.np_int[2] + o[2])
          IDEFIX_ERROR("Dump: the grid of the restart file does not match the current grid");
        long long c = 0;
        for (int k = data.beg[2]; k < data.end[2] + o[2]; k++)
          for (int j = data.beg[1]; j < data.end[1] + o[1]; j++)
            for (int i = data.beg[0]; i < data.end[0] + o[0]; i++) d.Ve(v, k, j, i) = dv[c++];
#endif
      } else if (field == "time") d.t = dv[0];
      else if (field == "dt") d.dt = dv[0];
      else if (field == "vtkLast") vtkLast = dv[0];
      else if (field == "dumpLast") dumpLast = dv[0];
      else if (field == "analysisLast") analysisLast = dv[0];
      else if (field == "vtkFileNumber") vtkFileNumber = iv[0];
      else if (field == "dumpFileNumber") dumpFileNumber = iv[0];
    }
    std::fclose(f);
    d.SyncToDevice();
    idfx::cout << "done." << std::endl;
    return true;
  }

  // ---- VTK files ---------------------------------------------------------------------------------------
  void WriteVtk(DataBlock &data) {
    using idfxhost::bigEndian;
    if (userDefVariablesEnabled) {
      if (!userDefVariablesFunc)
        IDEFIX_ERROR("Cannot output user-defined variables without enrollment of your user-defined variables function");
      userDefVariablesFunc(data, userDefVariables);
    }
    const std::string name = vtkDir + idfxhost::numbered("data", vtkFileNumber, "vtk");
    idfx::cout << "Vtk: Write file " << name << "..." << std::flush;
    const auto t0 = std::chrono::steady_clock::now();
    DataBlockHost d(data);
    d.SyncFromDevice();
    Grid &grid = *data.mygrid;
    const int nx[3] = {grid.np_int[0], grid.np_int[1], grid.np_int[2]};
    const int off[3] = {grid.np_tot[0] == 1 ? 0 : 1, grid.np_tot[1] == 1 ? 0 : 1, grid.np_tot[2] == 1 ? 0 : 1};
    std::vector<float> node[3], center[3];   // already big-endian
    for (int dir = 0; dir < 3; dir++) {
      const int g = grid.nghost[dir];
      for (int i = 0; i < nx[dir] + off[dir]; i++)
        node[dir].push_back(bigEndian(static_cast<float>(grid.np_tot[dir] == 1 ? grid.x[dir](i) : grid.xl[dir](i + g))));
      for (int i = 0; i < nx[dir]; i++)
        center[dir].push_back(grid.np_tot[dir] == 1 ? node[dir][i] : bigEndian(static_cast<float>(grid.x[dir](i + g))));
    }
    std::FILE *f = std::fopen(name.c_str(), "wb");
    if (!f) IDEFIX_ERROR("Unable to open file " + name);
    auto text = [&](const std::string &s) { std::fwrite(s.data(), 1, s.size(), f); };
    auto floats = [&](const std::vector<float> &v) { std::fwrite(v.data(), 4, v.size(), f); };
    const bool rectilinear = (GEOMETRY == CARTESIAN || GEOMETRY == CYLINDRICAL);
    text("# vtk DataFile Version 2.0\n");
    text(std::string("Idefix ") + IDEFIX_VERSION + " VTK Data\n");
    text("BINARY\n");
    text(rectilinear ? "DATASET RECTILINEAR_GRID\n" : "DATASET STRUCTURED_GRID\n");
    text("FIELD FieldData 9\n");
    text("GEOMETRY 1 1 int\n");
    const int32_t geo = bigEndian(static_cast<int32_t>(GEOMETRY));
    std::fwrite(&geo, 4, 1, f);
    text("\nPERIODICITY 1 3 int\n");
    for (int dir = 0; dir < 3; dir++) {
      const int32_t p = bigEndian(static_cast<int32_t>(grid.lbound[dir] == periodic));
      std::fwrite(&p, 4, 1, f);
    }
    text("\nTIME 1 1 float\n");
    const float tb = bigEndian(static_cast<float>(data.t));
    std::fwrite(&tb, 4, 1, f);
    for (int dir = 0; dir < 3; dir++) {
      text("\nX" + std::to_string(dir + 1) + "L_NATIVE_COORDINATES 1 " + std::to_string(nx[dir] + off[dir]) + " float\n");
      floats(node[dir]);
    }
    for (int dir = 0; dir < 3; dir++) {
      text("\nX" + std::to_string(dir + 1) + "C_NATIVE_COORDINATES 1 " + std::to_string(nx[dir]) + " float\n");
      floats(center[dir]);
    }
    text("\nDIMENSIONS " + std::to_string(nx[0] + off[0]) + " " + std::to_string(nx[1] + off[1]) + " " +
         std::to_string(nx[2] + off[2]) + "\n");
    if (rectilinear) {
      text("X_COORDINATES " + std::to_string(nx[0] + off[0]) + " float\n");
      floats(node[0]);
      text("\nY_COORDINATES " + std::to_string(nx[1] + off[1]) + " float\n");
      floats(node[1]);
      text("\nZ_COORDINATES " + std::to_string(nx[2] + off[2]) + " float\n");
      floats(node[2]);
    } else {
      // node positions converted to cartesian (vtk.cpp:180-260)
      const long long npts = static_cast<long long>(nx[0] + off[0]) * (nx[1] + off[1]) * (nx[2] + off[2]);
      text("POINTS " + std::to_string(npts) + " float\n");
      std::vector<float> pts;
      pts.reserve(3 * npts);
      for (int k = 0; k < nx[2] + off[2]; k++)
        for (int j = 0; j < nx[1] + off[1]; j++)
          for (int i = 0; i < nx[0] + off[0]; i++) {
            const double x1 = grid.np_tot[0] == 1 ? grid.x[0](i) : grid.xl[0](i + grid.nghost[0]);
            const double x2 = grid.np_tot[1] == 1 ? grid.x[1](j) : grid.xl[1](j + grid.nghost[1]);
            const double x3 = grid.np_tot[2] == 1 ? grid.x[2](k) : grid.xl[2](k + grid.nghost[2]);
            double px, py, pz;
            if (GEOMETRY == POLAR) {
              px = x1 * std::cos(x2);
              py = x1 * std::sin(x2);
              pz = x3;
            } else {   // spherical
              const double th = (DIMENSIONS == 1) ? M_PI / 2 : x2;
              px = x1 * std::sin(th) * std::cos(x3);
              py = x1 * std::sin(th) * std::sin(x3);
              pz = x1 * std::cos(th);
              if (DIMENSIONS <= 2) {
                py = pz;   // 2-D spherical runs are drawn in the meridional plane
                pz = 0.0;
              }
            }
            pts.push_back(bigEndian(static_cast<float>(px)));
            pts.push_back(bigEndian(static_cast<float>(py)));
            pts.push_back(bigEndian(static_cast<float>(pz)));
          }
      floats(pts);
    }
    text("\nCELL_DATA " + std::to_string(static_cast<long long>(nx[0]) * nx[1] * nx[2]) + "\n");
    std::map<std::string, std::function<double(int, int, int)>> scalars;
    const auto &names = data.hydro->VcName;
    for (int n = 0; n < static_cast<int>(names.size()); n++)
      scalars[names[n]] = [&d, n](int k, int j, int i) { return d.Vc(n, k, j, i); };
    for (auto &u : userDefVariables) {
      auto arr = u.second;
      scalars[u.first] = [arr](int k, int j, int i) { return arr(k, j, i); };
    }
    std::vector<float> buf(static_cast<size_t>(nx[0]) * nx[1] * nx[2]);
    for (auto &s : scalars) {
      size_t c = 0;
      for (int k = data.beg[2]; k < data.end[2]; k++)
        for (int j = data.beg[1]; j < data.end[1]; j++)
          for (int i = data.beg[0]; i < data.end[0]; i++) buf[c++] = bigEndian(static_cast<float>(s.second(k, j, i)));
      text("\nSCALARS " + s.first + " float\nLOOKUP_TABLE default\n");
      floats(buf);
    }
    std::fclose(f);
    vtkFileNumber++;
    idfx::cout << "done in " << std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() << " s."
               << std::endl;
  }

 private:
  AnalysisFunc analysisFunc{nullptr};
  UserDefVariablesFunc userDefVariablesFunc{nullptr};
  double elapsed{0.0};

  static void WriteField(std::FILE *f, const std::string &name, int32_t type, std::vector<int> dims, const void *p) {
    char nm[16] = {};
    std::strncpy(nm, name.c_str(), 15);
    std::fwrite(nm, 1, 16, f);
    const int32_t nd = static_cast<int32_t>(dims.size());
    std::fwrite(&type, 4, 1, f);
    std::fwrite(&nd, 4, 1, f);
    long long n = 1;
    for (int dd : dims) {
      const int32_t v = dd;
      std::fwrite(&v, 4, 1, f);
      n *= dd;
    }
    std::fwrite(p, type == 0 ? 8 : 4, n, f);
  }
  template <typename G>
  static std::vector<double> Interior(const DataBlockHost &d, G get, int o0, int o1, int o2) {
    std::vector<double> buf;
    buf.reserve(static_cast<size_t>(d.np_int[0] + o0) * (d.np_int[1] + o1) * (d.np_int[2] + o2));
    for (int k = d.beg[2]; k < d.end[2] + o2; k++)
      for (int j = d.beg[1]; j < d.end[1] + o1; j++)
        for (int i = d.beg[0]; i < d.end[0] + o0; i++) buf.push_back(get(k, j, i));
    return buf;
  }
};

#endif  // IDEFIX_B200_HOST_OUTPUT_HPP_
