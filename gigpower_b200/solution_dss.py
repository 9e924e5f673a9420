"""``SolutionDSS`` is OpenDSS itself behind gigpower's result accessors (solution_dss.py of the reference): it is
not part of the accelerated path and OpenDSS is not a dependency of this
package.  Importing the name fails with an explanation instead of an AttributeError somewhere later."""


def __getattr__(name):
    if name == 'SolutionDSS':
        raise ImportError('gigpower_b200 has no SolutionDSS: it would be OpenDSS itself, which this package neither '
                          'needs nor ships; use gigpower.solution_dss.SolutionDSS next to the accelerated '
                          'SolutionFBS / SolutionNR3 when an OpenDSS comparison is wanted')
    raise AttributeError(name)
