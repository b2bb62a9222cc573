"""Stand-in for the third-party `problog` package (absent from the image, not installable offline), used ONLY by
`bench.py --impl reference`: the reference's models/vael.py and utils/graph_semiring.py import two names from it
at module level (`problog.logic.{Term,Constant}`, `problog.evaluator.Semiring`).  Nothing here computes anything
on the timed path; it does not import vael_b200."""
