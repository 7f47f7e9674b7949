// CheckParameter.hpp — sanity checks before each shot (mirror of src/CheckParameter/CheckParameter.hpp:60-260).
#pragma once
#include "Acquisition.hpp"
#include "Modelparameter.hpp"

namespace KITGPI
{
    namespace CheckParameter
    {
        //! tells how the model dimensions and the interfaces were moved to fit the variable grid (CheckParameter.hpp:22-49)
        template <typename ValueType> void checkVariableGrid(Configuration::Configuration const &config, Acquisition::Coordinates<ValueType> const &modelCoordinates)
        {
            const IndexType NX = config.get<IndexType>("NX"), NY = config.get<IndexType>("NY"), NZ = config.get<IndexType>("NZ");
            const IndexType newNX = modelCoordinates.getNX(), newNY = modelCoordinates.getNY(), newNZ = modelCoordinates.getNZ();
            if (NX != newNX || NY != newNY || NZ != newNZ)
                HOST_PRINT("\nIn order to fit the variable grid, the model dimension had been altered from:" << NX << " X " << NY << " X " << NZ << " to " << newNX << " X " << newNY << " X "
                                                                                                            << newNZ << " \n")
            auto const &newInterfaces = modelCoordinates.getInterfaceVec();
            std::vector<IndexType> interface = Acquisition::readColumnFromFile(config.get<std::string>("gridConfigurationFilename"), 0);
            if (!interface.empty() && interface.at(0) == 0)
                interface.erase(interface.begin());
            else
                COMMON_THROWEXCEPTION("First interface must by at y=0 ")
            for (size_t i = 0; i < interface.size() && i + 1 < newInterfaces.size(); i++)
                if (interface[i] != newInterfaces[i + 1])
                    HOST_PRINT("In order to fit the variable grid, the interface Nr." << i + 1 << " has benn moved from Y=" << interface[i] << " to Y=" << newInterfaces[i + 1] << "\n\n")
        }

        //! Courant-Friedrichs-Lewy criterion.  The reference's h factors (7/6, 149/120, ...) are integer divisions and
        //! evaluate to 1 (CheckParameter.hpp:73-98), so the enforced bound is dt <= DH / (sqrt(D) vpMax) for every order.
        template <typename ValueType>
        void checkStabilityCriterion(ValueType dt, ValueType DH, ValueType vpMax, std::string dimension, IndexType spFDo, IndexType shotNumber = -1)
        {
            IndexType D;
            std::transform(dimension.begin(), dimension.end(), dimension.begin(), ::tolower);
            if (dimension == "2d")
                D = 2;
            else if (dimension == "3d")
                D = 3;
            else
                COMMON_THROWEXCEPTION("Unknown dimension")
            SCAI_ASSERT_ERROR(spFDo >= 2 && spFDo <= 12 && spFDo % 2 == 0, "Unknown spatial FD order")
            const ValueType h = 1;
            if (dt > DH / (h * std::sqrt((ValueType)D) * vpMax)) {
                HOST_PRINT("\nCourant-Friedrichs-Lewy-Criterion is not met" << (shotNumber >= 0 ? " for shot number: " + std::to_string(shotNumber) : std::string()) << "! \ndt is " << dt
                                                                            << " but should be less than DH/(h*sqrt(D)*vpMax=" << DH / (h * std::sqrt((ValueType)D) * vpMax) << "\n\n")
                COMMON_THROWEXCEPTION("\n\nCourant-Friedrichs-Lewy-Criterion is not met! \n\n")
            }
        }

        //! points per minimum wavelength (warning only, CheckParameter.hpp:122-157)
        template <typename ValueType> void checkNumericalDispersion(ValueType DH, ValueType vMin, ValueType fcMax, IndexType spFDo, IndexType shotNumber = -1)
        {
            const IndexType Ns[] = {12, 8, 7, 6, 5, 4};
            SCAI_ASSERT_ERROR(spFDo >= 2 && spFDo <= 12 && spFDo % 2 == 0, "Unknown spatial FD order")
            const IndexType N = Ns[spFDo / 2 - 1];
            if (DH > vMin / (2 * fcMax * N))
                HOST_PRINT("", "\nCriterion to avoid numerical dispersion is not met" << (shotNumber >= 0 ? " for shot number: " + std::to_string(shotNumber) : std::string()) << "! \nDH is "
                                                                                      << DH << " but should be less than vMin/(2*fcMax*N)=" << vMin / (2 * fcMax * N) << "\n\n")
        }

        template <typename ValueType>
        void checkNumericalArtefactsAndInstabilities(Configuration::Configuration const &config, std::vector<Acquisition::sourceSettings<ValueType>> const &sourceSettings,
                                                     Modelparameter::Modelparameter<ValueType> const &model, Acquisition::Coordinates<ValueType> const &modelCoordinates,
                                                     IndexType shotNumber = -1)
        {
            const std::string type = model.getEquationType();
            if (type == "elastic" || type == "viscoelastic") {
                auto const &vp = model.getVelocityP(), &vs = model.getVelocityS();
                for (size_t i = 0; i < vp.size(); i++)
                    SCAI_ASSERT_ERROR(vp[i] / vs[i] >= std::sqrt(2.0), "\n vp/vs (" << vp[i] << "/" << vs[i] << ") < sqrt(2.0) at index " << i << "\n\n")
            }
            ValueType fcMax = 0;
            for (auto const &s : sourceSettings)
                fcMax = std::max(fcMax, s.fc);
            checkStabilityCriterion<ValueType>(config.get<ValueType>("DT"), modelCoordinates.getDH(), model.getMaxVelocity(), config.get<std::string>("dimension"),
                                               config.get<IndexType>("spatialFDorder"), shotNumber);
            checkNumericalDispersion<ValueType>(modelCoordinates.getDH(), model.getMinVelocity(), fcMax, config.get<IndexType>("spatialFDorder"), shotNumber);
        }

        //! every source / receiver must lie inside the grid (CheckParameter.hpp:259-300)
        template <typename ValueType, typename Settings> void checkAcquisition(std::vector<Settings> const &settings, Acquisition::Coordinates<ValueType> const &modelCoordinates, const char *what)
        {
            for (auto const &s : settings) {
                auto c = s.getCoords();
                SCAI_ASSERT_ERROR(c.x >= 0 && c.x < modelCoordinates.getNX() && c.y >= 0 && c.y < modelCoordinates.getNY() && c.z >= 0 && c.z < modelCoordinates.getNZ(),
                                  what << " coordinate (" << c.x << "," << c.y << "," << c.z << ") is outside the model grid")
            }
        }
    }
}
