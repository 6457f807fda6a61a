// denovo_main.cpp — DenovoMain (M/DenovoMain.java:26-44): parse flags, print usage + exit 1 on a
// flag error, then DenovoRunner.execute.  ParseException is reported and the run ends (exit 0,
// like the reference's `return` after printStackTrace); anything else is a non-zero exit.
#include <cstdio>
#include <exception>

#include "denovo_host.hpp"

int main(int argc, char** argv) {
  denovo::CommandLine cmd;
  try {
    cmd.setArgs(argc, argv);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n\n%s", e.what(), denovo::CommandLine::usage().c_str());
    return 1;
  }
  if (cmd.help) {
    std::fputs(denovo::CommandLine::usage().c_str(), stdout);
    return 0;
  }
  try {
    denovo::run(cmd);
  } catch (const denovo::ParseException& e) {
    std::fprintf(stderr, "java.text.ParseException: %s\n", e.what());
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "Exception in thread \"main\": %s\n", e.what());
    return 2;
  }
  return 0;
}
