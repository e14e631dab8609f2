// input.hpp -- Input: idefix.ini reader + command line (reference: src/input.hpp:15-69 accessors,
// src/input.cpp:30-110 grammar, :113-190 command line, :193-245 ShowConfig log lines).
#include "idefix.hpp"
#ifndef IDEFIX_B200_HOST_INPUT_HPP_
#define IDEFIX_B200_HOST_INPUT_HPP_

class Input {
 public:
  using Params = std::vector<std::string>;
  using Block = std::map<std::string, Params>;
  std::map<std::string, Block> inputParameters;
  std::string inputFileName{"idefix.ini"};
  bool restartRequested{false};
  int restartFileNumber{-1};
  bool forceInitRequested{false};
  bool forceNoWrite{false};
  int maxCycles{-1};
  bool profile{false};
  int device{0};   // extension: -device n selects the CUDA device

  Input() = default;
  Input(int argc, char **argv) {
    ParseCommandLine(argc, argv);
    std::ifstream file(inputFileName);
    if (!file) IDEFIX_ERROR("Input: cannot open input file " + inputFileName);
    std::stringstream ss;
    ss << file.rdbuf();
    Parse(ss.str());
  }
  static Input FromString(const std::string &text) {
    Input in;
    in.inputFileName = "<string>";
    in.Parse(text);
    return in;
  }

  // [Block] headers, "name v0 v1 ..." entries, '#' starts a comment anywhere
  void Parse(const std::string &text) {
    std::istringstream all(text);
    std::string line, block;
    bool haveBlock = false;
    int nparams = 0;
    while (std::getline(all, line)) {
      const size_t hash = line.find('#');
      if (hash != std::string::npos) line = line.substr(0, hash);
      const size_t first = line.find_first_not_of(" \t\r");
      if (first == std::string::npos) continue;
      if (line[first] == '[') {
        const size_t end = line.find(']', first + 1);
        if (end == std::string::npos)
          IDEFIX_ERROR("Block name in file '" + inputFileName + "' not properly ended");
        if (haveBlock && nparams == 0) inputParameters[block]["!!empty!!"].push_back("!!empty!!");
        block = line.substr(first + 1, end - first - 1);
        inputParameters[block];
        haveBlock = true;
        nparams = 0;
        continue;
      }
      if (!haveBlock)
        IDEFIX_ERROR("Input file '" + inputFileName + "' must specify a block name before the first parameter");
      std::istringstream ls(line);
      std::string name, v;
      ls >> name;
      nparams++;
      Params &p = inputParameters[block][name];
      while (ls >> v) p.push_back(v);
    }
  }

  void ParseCommandLine(int argc, char **argv) {
    bool enableLogs = true;
    for (int i = 1; i < argc; i++) {
      const std::string a(argv[i]);
      auto needInt = [&](const char *what) {
        if (i + 1 >= argc || !std::isdigit(static_cast<unsigned char>(argv[i + 1][0])))
          IDEFIX_ERROR(std::string(what) + " requires an additional integer parameter");
        return std::stoi(std::string(argv[++i]));
      };
      if (a == "-i") {
        if (++i >= argc) IDEFIX_ERROR("You must specify -i filename where filename is the name of the input file.");
        inputFileName = argv[i];
      } else if (a == "-maxcycles") {
        maxCycles = needInt("-maxcycles");
        inputParameters["CommandLine"]["maxCycles"].push_back(std::to_string(maxCycles));
      } else if (a == "-restart") {
        std::string n = "-1";
        if (i + 1 < argc && std::isdigit(static_cast<unsigned char>(argv[i + 1][0]))) n = argv[++i];
        inputParameters["CommandLine"]["restart"].push_back(n);
        restartRequested = true;
        restartFileNumber = std::stoi(n);
      } else if (a == "-dec") {
        IDEFIX_ERROR("Domain decomposition option '-dec' only makes sense when MPI is enabled "
                     "(multi-GPU runs use one process per GPU through the C ABI's halo hooks)");
      } else if (a == "-force_init") {
        forceInitRequested = true;
      } else if (a == "-nowrite") {
        forceNoWrite = true;
        enableLogs = false;
      } else if (a == "-nolog") {
        enableLogs = false;
      } else if (a == "-profile") {
        profile = true;
      } else if (a == "-Werror") {
      } else if (a == "-device") {
        device = needInt("-device");
      } else if (a == "-version" || a == "-v") {
        idfx::cout << "Idefix version " << IDEFIX_VERSION << " (" << idfx_version() << ")" << std::endl;
        idfx::safeExit(0);
      } else {
        IDEFIX_ERROR("Unknown option " + a);
      }
    }
    if (enableLogs) idfx::cout.enableLogFile("idefix.0.log");
  }

  bool CheckBlock(const std::string &block) const { return inputParameters.count(block) > 0; }
  // number of values of the entry, -1 when absent (src/input.cpp:316-330)
  int CheckEntry(const std::string &block, const std::string &name) const {
    auto b = inputParameters.find(block);
    if (b == inputParameters.end()) return -1;
    auto p = b->second.find(name);
    if (p == b->second.end()) return -1;
    return static_cast<int>(p->second.size());
  }

  template <typename T>
  T Get(const std::string &block, const std::string &name, int num) {
    if (CheckEntry(block, name) <= num) {
      std::stringstream msg;
      msg << "Mandatory parameter [" << block << "]:" << name << "(" << num << ") cannot be found in the input file.";
      IDEFIX_ERROR(msg);
    }
    return Convert<T>(inputParameters[block][name][num], block, name, num);
  }

  // src/input.hpp:147-170: the default is stored through an ostream (6 significant digits) and read
  // back, so e.g. the default gamma 5/3 becomes 1.66667 -- reproduced because it changes results.
  template <typename T>
  T GetOrSet(const std::string &block, const std::string &name, int num, T def) {
    const int size = CheckEntry(block, name);
    if (size <= num) {
      if (size < 0 && num > 0) IDEFIX_ERROR("Entry [" + block + "]:" + name + " is not defined.");
      if (size + 1 < num) IDEFIX_ERROR("Entry [" + block + "]:" + name + " has too few parameters.");
      std::stringstream ss;
      if constexpr (std::is_same<T, bool>::value) ss << std::boolalpha << def;
      else ss << def;
      inputParameters[block][name].push_back(ss.str());
    }
    return Get<T>(block, name, num);
  }

  void ShowConfig() {
    const char *bar = "-----------------------------------------------------------------------------";
    idfx::cout << bar << std::endl
               << "Input Parameters using input file " << inputFileName << ":" << std::endl
               << bar << std::endl;
    for (auto &b : inputParameters) {
      idfx::cout << "[" << b.first << "]" << std::endl;
      for (auto &p : b.second) {
        idfx::cout << "\t" << p.first << "\t";
        for (auto &v : p.second) idfx::cout << "\t" << v;
        idfx::cout << std::endl;
      }
    }
    idfx::cout << bar << std::endl << bar << std::endl;
    idfx::cout << "Input: execution space: " << idfx_version() << std::endl << bar << std::endl;
    idfx::cout << "Input: Compiled with DOUBLE PRECISION arithmetic." << std::endl;
    idfx::cout << "Input: DIMENSIONS=" << DIMENSIONS << "." << std::endl;
    idfx::cout << "Input: COMPONENTS=" << COMPONENTS << "." << std::endl;
  }

 private:
  template <typename T>
  T Convert(const std::string &raw, const std::string &block, const std::string &name, int num) {
    try {
      if constexpr (std::is_same<T, std::string>::value) {
        return raw;
      } else if constexpr (std::is_same<T, bool>::value) {
        std::string low(raw);
        std::transform(low.begin(), low.end(), low.begin(), ::tolower);
        if (low == "yes" || low == "true" || low == "1") return true;
        if (low == "no" || low == "false" || low == "0") return false;
        throw std::invalid_argument(raw);
      } else if constexpr (std::is_integral<T>::value) {
        // "1e3"-style integers are accepted like the reference does (through a double)
        return static_cast<T>(std::round(std::stod(raw)));
      } else {
        return static_cast<T>(std::stod(raw));
      }
    } catch (const std::exception &) {
      std::stringstream msg;
      msg << "Cannot convert [" << block << "]:" << name << "(" << num << ") = '" << raw << "'";
      IDEFIX_ERROR(msg);
    }
    return T();
  }
};

#endif  // IDEFIX_B200_HOST_INPUT_HPP_
