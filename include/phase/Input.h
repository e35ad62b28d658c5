// Input.h -- the four .info trees of a case (reference: src/System/Input.{h,cpp}).
#ifndef PHASE_B200_INPUT_H
#define PHASE_B200_INPUT_H
#include "PropertyTree.h"

class Input {
public:
  Input(const std::string &caseDirectory = "case", const std::string &outputPath = "solution")
      : caseDirectory(caseDirectory), outputPath(outputPath) {}
  // S/Input.cpp:8-15 (initialConditions / postProcessing are optional here)
  void parseInputFile() {
    caseInput_ = read("case.info");
    boundaryInput_ = read("boundaries.info");
    try { initialConditionInput_ = read("initialConditions.info"); } catch (const Exception &) {}
    try { postProcessingInput_ = read("postProcessing.info"); } catch (const Exception &) {}
  }
  std::string caseDirectory, outputPath;
  const boost::property_tree::ptree &caseInput() const { return caseInput_; }
  const boost::property_tree::ptree &boundaryInput() const { return boundaryInput_; }
  const boost::property_tree::ptree &initialConditionInput() const { return initialConditionInput_; }
  const boost::property_tree::ptree &postProcessingInput() const { return postProcessingInput_; }
  boost::property_tree::ptree &caseInput() { return caseInput_; }
  boost::property_tree::ptree &boundaryInput() { return boundaryInput_; }
  boost::property_tree::ptree read(const std::string &filename) const {
    return phase::PropertyTree::readInfo(caseDirectory + "/" + filename);
  }

private:
  boost::property_tree::ptree caseInput_, boundaryInput_, initialConditionInput_, postProcessingInput_;
};
#endif
